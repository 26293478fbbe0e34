"""fhe_precompiles_b200: B200-native BFV precompile engine behind the fhe_precompiles surface.

`FHE` / `FheApp` -- the reference's precompile methods (bytes in, bytes out) over the CUDA C-ABI library
`pack`           -- pack.rs framing helpers
`device`         -- device-resident batched entry points on torch CUDA tensors (benchmarks, tests)
"""
from .fhe import FHE, FheApp  # noqa: F401
from .pack import FheError  # noqa: F401
