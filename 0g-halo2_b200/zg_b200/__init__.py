"""zg_b200 -- B200-native halo2/KZG proving backend for BN254 (host-side Python mirror).

`lib.Context` wraps the C ABI of libzg_b200.so (include/zg_b200.h); `bn254_host` holds the
field constants and byte-layout helpers the host side needs.  The CUDA library is the
product: nothing in this package computes prover arithmetic on the CPU."""
from . import bn254_host  # noqa: F401
from .lib import Context, ZgError, load_library, LIB_PATH, BASIS_MONOMIAL, BASIS_LAGRANGE  # noqa: F401
