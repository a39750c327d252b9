"""Host-side mirrors of the reference's op packages, backed by the sm_100a C-ABI library."""
