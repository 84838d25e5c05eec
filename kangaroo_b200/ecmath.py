"""Host-side secp256k1 glue for the stand-alone solver (NOT on the jump path): public-key parsing, the handful of
scalar multiplications needed to set up a search (jump-table points, start*G) and to verify a found key.
Pure Python big ints; a few hundred point operations per search at most.  The reference does the same work with
SECPK1/SECP256K1.cpp on the CPU (ParsePublicKeyHex :140-230, ComputePublicKey :59-87, AddDirect :238-262)."""
P = 2**256 - 0x1000003D1
N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
G = (0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798,
     0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8)


def add(a, b):
    """affine add; None is the point at infinity"""
    if a is None:
        return b
    if b is None:
        return a
    if a[0] == b[0]:
        if (a[1] + b[1]) % P == 0:
            return None
        s = 3 * a[0] * a[0] * pow(2 * a[1], -1, P) % P
    else:
        s = (b[1] - a[1]) * pow(b[0] - a[0], -1, P) % P
    x = (s * s - a[0] - b[0]) % P
    return x, (s * (a[0] - x) - a[1]) % P


def neg(a):
    return None if a is None else (a[0], (-a[1]) % P)


def mul(k, pt=G):
    k %= N
    acc = None
    while k:
        if k & 1:
            acc = add(acc, pt)
        pt = add(pt, pt)
        k >>= 1
    return acc


def parse_pubkey(hexstr):
    """compressed (02/03 + x) or uncompressed (04 + x + y) hex -> (x, y)"""
    h = hexstr.strip()
    if h[:2] in ("02", "03") and len(h) == 66:
        x = int(h[2:], 16)
        y = pow((x * x * x + 7) % P, (P + 1) // 4, P)
        if (y & 1) != (int(h[:2], 16) & 1):
            y = P - y
        assert (y * y - x * x * x - 7) % P == 0, "not on the curve"
        return x, y
    if h[:2] == "04" and len(h) == 130:
        return int(h[2:66], 16), int(h[66:], 16)
    raise ValueError("bad public key: " + hexstr)


def parse_config(path):
    """Kangaroo::ParseConfigFile (Kangaroo.cpp:84-144): start, end, then public keys, all hex."""
    lines = [ln.strip() for ln in open(path) if ln.strip()]
    return int(lines[0], 16), int(lines[1], 16), [parse_pubkey(ln) for ln in lines[2:]]
