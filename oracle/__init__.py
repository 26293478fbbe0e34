"""TEST INFRASTRUCTURE ONLY: CPU oracle for the BFV precompile hot path.

`oracle.bfv`     -- ctypes binding of liboracle.so (C restatement of the SEAL 4.0 arithmetic)
`oracle.formats` -- pure-Python restatement of the wire formats (pack.rs + bincode + SEAL + zstd)

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this package.
"""
