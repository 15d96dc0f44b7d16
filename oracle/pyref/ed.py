"""ORACLE (test infrastructure, never shipped) -- GF(2^255-19), Edwards25519 and ristretto255.

Big-int restatement of the arithmetic that curve25519-dalek 3.2.0 performs for the
reference (pinned at /root/reference/Cargo.lock:155-157; the crate source is NOT vendored
under /root/reference, so this file follows the published algorithm, RFC 9496).

Reference call sites this stands in for:
  * CompressedRistretto::{from_slice,as_bytes}   /root/reference/src/lalrpop/assignment_parser.rs:137,205
  * RistrettoPoint arithmetic behind Prover::commit /root/reference/src/gadget.rs:32

Pinned by: RFC 9496 appendix A vectors (multiples of the generator, bad encodings,
hash-to-group) -- see tests/test_oracle_anchors.py.  Parity with dalek itself: UNPINNED
(no Rust toolchain in the build image).
"""

P = 2**255 - 19
D = (-121665 * pow(121666, P - 2, P)) % P
SQRT_M1 = 19681161376707505956807079304988542015446066515923890162744021073123829784752
SQRT_AD_MINUS_ONE = 25063068953384623474111414158702152701244531502492656460079210482610430750235
INVSQRT_A_MINUS_D = 54469307008909316920995813868745141605393597292927456921205312896311721017578
ONE_MINUS_D_SQ = (1 - D * D) % P
D_MINUS_ONE_SQ = ((D - 1) * (D - 1)) % P


def is_neg(x):
    return (x % P) & 1


def ct_abs(x):
    x %= P
    return (P - x) % P if x & 1 else x


def sqrt_ratio_m1(u, v):
    """RFC 9496 4.2 SQRT_RATIO_M1 -> (was_square, r)."""
    u %= P
    v %= P
    v3 = v * v % P * v % P
    v7 = v3 * v3 % P * v % P
    r = (u * v3) % P * pow(u * v7 % P, (P - 5) // 8, P) % P
    check = v * r % P * r % P
    correct = check == u
    flipped = check == (P - u) % P
    flipped_i = check == (P - u) * SQRT_M1 % P
    if flipped or flipped_i:
        r = r * SQRT_M1 % P
    return (correct or flipped), ct_abs(r)


class Point:
    """Extended twisted-Edwards coordinates (X:Y:Z:T), a=-1."""

    __slots__ = ("X", "Y", "Z", "T")

    def __init__(self, X, Y, Z, T):
        self.X, self.Y, self.Z, self.T = X % P, Y % P, Z % P, T % P

    @staticmethod
    def identity():
        return Point(0, 1, 1, 0)

    def __add__(self, o):
        # add-2008-hwcd-3 (a=-1), unified
        A = (self.Y - self.X) * (o.Y - o.X) % P
        B = (self.Y + self.X) * (o.Y + o.X) % P
        C = self.T * 2 * D % P * o.T % P
        Dd = self.Z * 2 * o.Z % P
        E, F, G, H = B - A, Dd - C, Dd + C, B + A
        return Point(E * F, G * H, F * G, E * H)

    def __neg__(self):
        return Point(-self.X, self.Y, self.Z, -self.T)

    def __sub__(self, o):
        return self + (-o)

    def double(self):
        return self + self

    def __mul__(self, k):
        k = int(k)
        if k < 0:
            return (-self) * (-k)
        acc = Point.identity()
        base = self
        while k:
            if k & 1:
                acc = acc + base
            base = base + base
            k >>= 1
        return acc

    __rmul__ = __mul__

    def affine(self):
        zi = pow(self.Z, P - 2, P)
        return self.X * zi % P, self.Y * zi % P

    def __eq__(self, o):
        # ristretto equality (RFC 9496 4.3.3)
        return (self.X * o.Y - self.Y * o.X) % P == 0 or (self.Y * o.Y - self.X * o.X) % P == 0

    def is_identity(self):
        return self == Point.identity()

    def compress(self):
        """RFC 9496 4.3.2 Encode."""
        x0, y0, z0, t0 = self.X, self.Y, self.Z, self.T
        u1 = (z0 + y0) * (z0 - y0) % P
        u2 = x0 * y0 % P
        _, invsqrt = sqrt_ratio_m1(1, u1 * u2 % P * u2 % P)
        den1 = invsqrt * u1 % P
        den2 = invsqrt * u2 % P
        z_inv = den1 * den2 % P * t0 % P
        ix0 = x0 * SQRT_M1 % P
        iy0 = y0 * SQRT_M1 % P
        ench = den1 * INVSQRT_A_MINUS_D % P
        if is_neg(t0 * z_inv):
            x, y, den_inv = iy0, ix0, ench
        else:
            x, y, den_inv = x0, y0, den2
        if is_neg(x * z_inv):
            y = (P - y) % P
        s = ct_abs(den_inv * (z0 - y))
        return s.to_bytes(32, "little")


def decompress(b):
    """RFC 9496 4.3.1 Decode; returns None for an invalid encoding."""
    if len(b) != 32:
        return None
    s = int.from_bytes(b, "little")
    if s >= P or (s & 1):
        return None
    ss = s * s % P
    u1 = (1 - ss) % P
    u2 = (1 + ss) % P
    u2_sqr = u2 * u2 % P
    v = (-(D * u1 % P * u1) - u2_sqr) % P
    was_square, invsqrt = sqrt_ratio_m1(1, v * u2_sqr % P)
    den_x = invsqrt * u2 % P
    den_y = invsqrt * den_x % P * v % P
    x = ct_abs(2 * s * den_x)
    y = u1 * den_y % P
    t = x * y % P
    if (not was_square) or is_neg(t) or y == 0:
        return None
    return Point(x, y, 1, t)


def elligator(t):
    """RFC 9496 4.3.4 MAP."""
    t %= P
    r = SQRT_M1 * t % P * t % P
    u = (r + 1) * ONE_MINUS_D_SQ % P
    v = (-1 - r * D) * (r + D) % P
    was_square, s = sqrt_ratio_m1(u, v)
    s_prime = (P - ct_abs(s * t)) % P
    if was_square:
        c = P - 1
    else:
        s, c = s_prime, r
    N = (c * (r - 1) % P * D_MINUS_ONE_SQ - v) % P
    w0 = 2 * s * v % P
    w1 = N * SQRT_AD_MINUS_ONE % P
    w2 = (1 - s * s) % P
    w3 = (1 + s * s) % P
    return Point(w0 * w3, w2 * w1, w1 * w3, w0 * w2)


def from_uniform_bytes(b):
    """RistrettoPoint::from_uniform_bytes: two Elligator maps, top bit of each half masked."""
    assert len(b) == 64
    mask = (1 << 255) - 1
    r0 = int.from_bytes(b[:32], "little") & mask
    r1 = int.from_bytes(b[32:], "little") & mask
    return elligator(r0) + elligator(r1)


BASEPOINT_COMPRESSED = bytes.fromhex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76")
BASEPOINT = decompress(BASEPOINT_COMPRESSED)


def msm(scalars, points):
    """Reference-semantics multiscalar mul (any algorithm gives the same group element)."""
    acc = Point.identity()
    for s, pt in zip(scalars, points):
        acc = acc + pt * int(s)
    return acc
