"""tfhe.jl_b200 — B200-native gate-bootstrapping engine behind the TFHE.jl API surface.

The directory name carries a dot, so import it through the repo-root shim ``tfhe_jl_b200``
(``import tfhe_jl_b200 as tfhe``), which registers this directory as a regular package.

Only what the hot path needs lives here: ``csrc/`` (sm_100a kernels + C ABI),
``_cabi.py`` (ctypes binding of the C ABI), ``api.py`` (host mirror of src/TFHE.jl:24-61) and
``julia/TFHEB200.jl`` (the ``ccall`` wrapper for real Julia hosts).
"""
from . import _cabi
from ._cabi import Context, TFHEB200Error, build, device_count, lib
from .api import (CloudKey, CloudKeyPart, DeviceLweBatch, LweSample, constant_dev, gate_dev, MKCloudKey, MKLweSample, SchemeParameters, SecretKey,
                  SharedKey, decrypt, encrypt, gate_and, gate_andny, gate_andyn, gate_constant, gate_mux,
                  gate_nand, gate_nor, gate_not, gate_or, gate_orny, gate_oryn, gate_xnor, gate_xor,
                  make_key_pair, mk_decrypt, mk_encrypt, mk_gate_nand, mktfhe_parameters_2party,
                  mktfhe_parameters_4party, mktfhe_parameters_8party, tfhe_parameters_80, tfhe_parameters_128)

__all__ = [n for n in dir() if not n.startswith("_")]

from .circuit import Circuit, Wire, adder_circuit, minimum_circuit  # noqa: E402,F401
from .api import (load_ciphertext, load_cloud_key, load_secret_key, save_ciphertext, save_cloud_key,  # noqa: E402,F401
                  save_secret_key)
