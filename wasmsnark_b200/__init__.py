"""wasmsnark_b200 -- B200 (sm_100a) replacement for the proving path of iden3/wasmsnark.

Host side: a thin ctypes mirror of the reference's `Bn128` host class over the C ABI of
libwasmsnark_b200.so (include/wasmsnark_b200.h).  All proving-path arithmetic runs in hand-written
CUDA kernels with no CPU fallback; only the Groth16 verifier (four pairings) is host code.
"""
from ._lib import load_library, LibraryNotBuilt  # noqa: F401
from .bn128 import Bn128, MultiBn128, ProvingKey, build_bn128, build_bn128_multi, groth16GenProof, genZKSnarkProof, groth16Verify, pairing, proof_json  # noqa: F401
