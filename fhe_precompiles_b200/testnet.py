"""Parameters of the first testnet: mirror of /root/reference/src/testnet.rs:8-14."""
LATTICE_DIMENSION = 4096
COEFF_MODULUS = (0xFFFFEE001, 0xFFFFC4001, 0x1FFFFE0001)
PLAIN_MODULUS = 4096
SCHEME = "bfv"
SECURITY_LEVEL = "TC128"
