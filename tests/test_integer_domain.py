"""The integer-domain form of the BEHZ base conversions (k_ext_conv / k_floor_sk) equals SEAL's step-by-step form
(fastbconv_m_tilde + sm_mrq, fast_floor + fastbconv_sk) -- CPU check of the algebra the kernels rely on."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load():
    spec = importlib.util.spec_from_file_location("check_integer_domain", os.path.join(ROOT, "scripts", "check_integer_domain.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_base_extension_integer_domain():
    _load().check_ext(20000)


def test_fast_floor_sk_integer_domain():
    _load().check_floor(20000)
