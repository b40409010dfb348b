"""EncryptedVector: a batch of Paillier ciphertexts resident in GPU memory.

The reference has no vector type: callers loop over scalars (examples/federated_learning_with_encryption.py
:122-133, phe/tests/math_test.py:44-58).  This is the batched form of exactly those loops -- one kernel
launch per vector operation, ciphertexts stay in HBM as [B, c_limbs] uint32 limb matrices and are
only converted to Python ints on request.  Semantics per element are those of EncryptedNumber
(exponent alignment before add, phe/paillier.py:695-700; lazy obfuscation, :565-566).
"""
import operator
import os

import numpy as np

from .encoding import EncodedNumber
from .engine import ints_to_limbs, limbs_to_ints, limbs_to_decimal_dev, decimal_to_limbs_dev, decimal_width


def _torch():
    import torch
    return torch


def _to_dev(arr, ctx):
    """numpy uint32 limb matrix -> torch.int32 tensor on the context's GPU (host tensor only for the test-only
    simulation build, whose "device" pointers are host pointers)."""
    torch = _torch()
    t = torch.from_numpy(np.ascontiguousarray(arr).view(np.int32).copy())
    return t if ctx.eng.simulated else t.to("cuda:%d" % ctx.device, non_blocking=False)


def _to_host(t):
    return t.cpu().numpy().view(np.uint32)


def _stream(ctx):
    """The CUDA stream the engine kernels of this call are launched on: torch's CURRENT stream of the context's
    device, so that they are ordered with the surrounding torch work (allocations, fills, index ops, .cpu()) also
    inside ``with torch.cuda.stream(s)``.  The engine keeps its scratch memory per (context, stream)."""
    if ctx.eng.simulated:
        return None
    return int(_torch().cuda.current_stream(ctx.device).cuda_stream)


def _rows_for(ctx_limbs, t):
    """Zero-pad (or trim all-zero columns of) a device limb matrix to `ctx_limbs` columns."""
    have = int(t.shape[1])
    if have == ctx_limbs:
        return t
    torch = _torch()
    if have < ctx_limbs:
        return torch.nn.functional.pad(t, (0, ctx_limbs - have)).contiguous()
    return t[:, :ctx_limbs].contiguous()


def limbs_to_decimal_strings(ctx, d_limbs):
    """Device limb matrix -> list of decimal strings (str(int) of every row), converted on the GPU."""
    torch = _torch()
    count, limbs = int(d_limbs.shape[0]), int(d_limbs.shape[1])
    if not count:
        return []
    width = decimal_width(limbs, ctx.eng)
    d_text = torch.empty((count, width), dtype=torch.uint8, device=d_limbs.device)
    limbs_to_decimal_dev(d_limbs, limbs, d_text, count, ctx.device, stream=_stream(ctx), engine=ctx.eng)
    raw = d_text.cpu().numpy().tobytes()
    return [(raw[i * width:(i + 1) * width].lstrip(b"0") or b"0").decode("ascii") for i in range(count)]


def decimal_strings_to_limbs(ctx, strings, limbs):
    """List of decimal strings -> device limb matrix [len, limbs] (int() of every string, on the GPU).
    ValueError for anything that is not a non-negative decimal integer fitting `limbs` limbs."""
    torch = _torch()
    count = len(strings)
    width = max([len(t) for t in strings] + [1])
    raw = b"".join(t.encode("ascii").rjust(width, b"0") if isinstance(t, str) else bytes(t).rjust(width, b"0") for t in strings)
    text = np.frombuffer(raw, dtype=np.uint8).reshape(count, width)
    t = torch.from_numpy(text.copy())
    d_text = t if ctx.eng.simulated else t.to("cuda:%d" % ctx.device)
    d_limbs = torch.empty((count, limbs), dtype=torch.int32, device=d_text.device)
    d_status = torch.zeros((count,), dtype=torch.int32, device=d_text.device)
    if count:
        decimal_to_limbs_dev(d_text, width, d_limbs, limbs, d_status, count, ctx.device, stream=_stream(ctx), engine=ctx.eng)
        bad = torch.nonzero(d_status).flatten()
        if bad.numel():
            i = int(bad[0])
            raise ValueError("value %d is not a decimal integer below 2**%d" % (i, 32 * limbs))
    return d_limbs


def random_r_values(n, count):
    """count values uniform in [1, n) from os.urandom (64 surplus bits: bias < 2^-64), the batched
    counterpart of PaillierPublicKey.get_random_lt_n (phe/paillier.py:141-143)."""
    nbytes = (n.bit_length() + 7) // 8 + 8
    raw = os.urandom(nbytes * count)
    span = n - 1
    return [1 + int.from_bytes(raw[i * nbytes:(i + 1) * nbytes], "little") % span for i in range(count)]


# ------------------------------------------------------------------------------------------------------
# Vectorised EncodedNumber.encode / decode: numpy fast paths that give exactly
# the values phe/encoding.py:110-233 computes element by element, with a per-element fallback for
# everything outside the fast path (huge mantissas, precision=..., non-finite values, ints beyond 63 bits).
def _limbs_from_signed(int_rep, n, ln):
    """[B] int64 signed integers -> [B, ln] uint32 limbs of (int_rep mod n), assuming |int_rep| < 2^63 < n."""
    mag = np.abs(int_rep).astype(np.uint64)
    neg = int_rep < 0
    out = np.zeros((int_rep.shape[0], ln), dtype=np.uint32)
    out[:, 0] = (mag & np.uint64(0xffffffff)).astype(np.uint32)
    if ln > 1:
        out[:, 1] = (mag >> np.uint64(32)).astype(np.uint32)
    if neg.any():
        nl = ints_to_limbs([n], ln)[0]
        n_low = np.uint64(int(nl[0]) | ((int(nl[1]) << 32) if ln > 1 else 0))
        m = mag[neg]
        low = n_low - m                                   # wraps modulo 2^64
        borrow = m > n_low
        rows = np.tile(nl, (m.shape[0], 1))
        rows[:, 0] = (low & np.uint64(0xffffffff)).astype(np.uint32)
        if ln > 1:
            rows[:, 1] = (low >> np.uint64(32)).astype(np.uint32)
        j = 2
        while j < ln and borrow.any():
            cur = rows[:, j]
            rows[:, j] = np.where(borrow, cur - np.uint32(1), cur)
            borrow = borrow & (cur == 0)
            j += 1
        out[neg] = rows
    return out


def encode_batch(public_key, values, precision=None, max_exponent=None):
    """Encode a sequence like EncodedNumber.encode does element by element.
    Returns (limbs [B, n_limbs] uint32 of the encodings, exponents [B] int64)."""
    ctx = public_key.engine_context()
    ln = ctx.n_limbs
    arr = None
    if precision is None and public_key.n.bit_length() > 80:
        if isinstance(values, np.ndarray) and values.dtype in (np.float64, np.int64, np.int32):
            arr = values
        elif len(values) and all(type(v) is float for v in values):
            arr = np.asarray(values, dtype=np.float64)
        elif len(values) and all(type(v) is int and -2 ** 62 < v < 2 ** 62 for v in values):
            arr = np.asarray(values, dtype=np.int64)
    if arr is not None and arr.dtype == np.float64 and np.isfinite(arr).all():
        m, e = np.frexp(arr)
        mant = np.ldexp(m, 53).astype(np.int64)                       # exact 53-bit signed mantissa
        lsb = e.astype(np.int64) - 53                                 # weight of the last mantissa bit
        exps = np.floor_divide(lsb, 4)                                # floor(lsb / log2(16)), phe/encoding.py:167-174
        if max_exponent is not None:
            exps = np.minimum(exps, np.asarray(max_exponent, dtype=np.int64))
        shift = lsb - 4 * exps
        if (shift <= 9).all():                                        # |int_rep| < 2^62
            int_rep = np.left_shift(mant, shift)
            return _limbs_from_signed(int_rep, public_key.n, ln), exps
    elif arr is not None and arr.dtype != np.float64:
        exps = np.zeros(arr.shape[0], dtype=np.int64)
        if max_exponent is None or (np.asarray(max_exponent) >= 0).all():
            return _limbs_from_signed(arr.astype(np.int64), public_key.n, ln), exps
    if isinstance(max_exponent, (list, tuple, np.ndarray)):
        mex = [int(x) for x in max_exponent]
    else:
        mex = [max_exponent] * len(values)
    vals = values.tolist() if isinstance(values, np.ndarray) else values
    encs = [v if isinstance(v, EncodedNumber) else EncodedNumber.encode(public_key, v, precision, mx)
            for v, mx in zip(vals, mex)]
    return (ints_to_limbs([x.encoding for x in encs], ln), np.array([x.exponent for x in encs], dtype=np.int64))


def decode_batch(public_key, limbs, exponents):
    """Decode plaintext limbs [B, n_limbs] with exponents [B] like EncodedNumber.decode; returns a list."""
    n = public_key.n
    ln = limbs.shape[1]
    count = limbs.shape[0]
    out = [None] * count
    exps = np.asarray(exponents, dtype=np.int64)
    small_pos = ~limbs[:, 2:].any(axis=1) if ln > 2 else np.ones(count, dtype=bool)
    low = limbs[:, 0].astype(np.uint64) | (limbs[:, 1].astype(np.uint64) << np.uint64(32)) if ln > 1 else limbs[:, 0].astype(np.uint64)
    # negatives are stored as n - |x| (phe/encoding.py:217-219): |x| = n - enc fits 64 bits exactly when the limbs above
    # the low pair equal those of n (no borrow out of the low pair) or of n - 2^64 (borrow)
    nl = ints_to_limbs([n], ln)[0]
    neg_small = np.zeros(count, dtype=bool)
    neg_mag = np.zeros(count, dtype=np.uint64)
    rest = np.nonzero(~small_pos)[0]
    if len(rest) and ln > 2 and n >> 64:
        n_low = int(n & (2 ** 64 - 1))
        hi0 = nl[2:]
        hi1 = ints_to_limbs([(n >> 64) - 1], ln - 2)[0]
        sub = limbs[rest]
        lo = low[rest]
        no_borrow = lo <= np.uint64(n_low)
        same0 = (sub[:, 2:] == hi0).all(axis=1)
        same1 = (sub[:, 2:] == hi1).all(axis=1)
        ok = np.where(no_borrow, same0, same1)
        neg_small[rest] = ok
        neg_mag[rest] = np.uint64(n_low) - lo              # modulo 2^64: the borrow is what `same1` accounts for
    fast = (small_pos | neg_small) & (exps < 0) & (exps > -250) & (public_key.max_int > 2 ** 64)
    if fast.any():
        mag = np.where(small_pos, low, neg_mag)
        val = np.ldexp(mag.astype(np.float64), (4 * exps).astype(np.int32)) if EncodedNumber.BASE == 16 else None
        val = np.where(small_pos, val, -val)
        if fast.all():
            out = val.tolist()
        else:
            for i, v in zip(np.nonzero(fast)[0].tolist(), val[fast].tolist()):
                out[i] = v
    slow = np.nonzero(~fast)[0]
    if len(slow):
        encs = limbs_to_ints(limbs[slow])
        for i, enc in zip(slow, encs):
            out[i] = EncodedNumber(public_key, enc, int(exps[i])).decode()
    return out


class EncryptedVector(object):
    __array_ufunc__ = None          # numpy_array + vector / numpy_array * vector defer to __radd__ / __rmul__

    def __init__(self, public_key, limbs, exponents, obfuscated=False):
        self.public_key = public_key
        self.limbs = limbs                                  # torch.int32 [B, c_limbs] on the key's device
        self.exponents = np.asarray(exponents, dtype=np.int64)
        self._obfuscated = obfuscated

    # ------------------------------------------------------------------ construction
    @classmethod
    def encrypt(cls, public_key, values, precision=None, r_values=None):
        """Encode every value (EncodedNumber.encode) and encrypt the batch in one launch.  With
        r_values None each element gets a fresh random r and is therefore already obfuscated."""
        if len(values) and any(isinstance(v, EncodedNumber) for v in (values if not isinstance(values, np.ndarray) else [])):
            encs = [v if isinstance(v, EncodedNumber) else EncodedNumber.encode(public_key, v, precision) for v in values]
            return cls.encrypt_encoded(public_key, [e.encoding for e in encs], [e.exponent for e in encs], r_values)
        limbs, exps = encode_batch(public_key, values, precision)
        return cls._encrypt_limbs(public_key, limbs, exps, r_values)

    @classmethod
    def encrypt_encoded(cls, public_key, encodings, exponents, r_values=None):
        ctx = public_key.engine_context()
        return cls._encrypt_limbs(public_key, ints_to_limbs([e % public_key.n for e in encodings], ctx.n_limbs), exponents, r_values)

    @classmethod
    def _encrypt_limbs(cls, public_key, m_limbs, exponents, r_values=None):
        ctx = public_key.engine_context()
        count = int(m_limbs.shape[0])
        obf = r_values is None
        torch = _torch()
        d_m = _to_dev(m_limbs, ctx)
        if r_values is None:
            # fresh obfuscators drawn on the device from a ChaCha20 stream keyed by os.urandom (pai_rng.cuh)
            d_r = torch.empty((count, ctx.n_limbs), dtype=torch.int32, device=d_m.device)
            if count:
                ctx.random_lt_n_dev(d_r, count, stream=_stream(ctx))
        else:
            d_r = _to_dev(ints_to_limbs(list(r_values), ctx.n_limbs), ctx)
        d_c = torch.empty((count, ctx.c_limbs), dtype=torch.int32, device=d_m.device)
        if count:
            ctx.encrypt_dev(d_m, d_r, d_c, count, stream=_stream(ctx))
        return cls(public_key, d_c, exponents, obfuscated=obf)

    @classmethod
    def from_encrypted_numbers(cls, numbers):
        pk = numbers[0].public_key
        ctx = pk.engine_context()
        limbs = _to_dev(ints_to_limbs([x.ciphertext(be_secure=False) % pk.nsquare for x in numbers], ctx.c_limbs), ctx)
        return cls(pk, limbs, [x.exponent for x in numbers])

    # ------------------------------------------------------------------ access
    def __len__(self):
        return int(self.limbs.shape[0])

    def ciphertexts(self, be_secure=True):
        if be_secure and not self._obfuscated:
            self.obfuscate()
        return limbs_to_ints(_to_host(self.limbs))

    def to_encrypted_numbers(self, be_secure=False):
        from .paillier import EncryptedNumber
        out = []
        for c, e in zip(self.ciphertexts(be_secure), self.exponents.tolist()):
            x = EncryptedNumber(self.public_key, c, int(e))
            if self._obfuscated:
                x._mark_obfuscated()
            out.append(x)
        return out

    def __getitem__(self, i):
        """v[int] -> EncryptedNumber (only that row leaves the device); v[slice / index array / mask] -> EncryptedVector."""
        import numbers
        if isinstance(i, numbers.Integral) and not isinstance(i, (bool, np.bool_)):
            from .paillier import EncryptedNumber
            k = operator.index(i)
            if k < 0:
                k += len(self)
            if not 0 <= k < len(self):
                raise IndexError("EncryptedVector index out of range")
            x = EncryptedNumber(self.public_key, limbs_to_ints(_to_host(self.limbs[k:k + 1]))[0], int(self.exponents[k]))
            if self._obfuscated:
                x._mark_obfuscated()
            return x
        if isinstance(i, np.ndarray):
            i = _torch().from_numpy(i).to(self.limbs.device)
            exps = self.exponents[i.cpu().numpy()]
        else:
            exps = self.exponents[i]
        return EncryptedVector(self.public_key, self.limbs[i].contiguous(), exps, self._obfuscated)

    def obfuscate(self):
        """c_i <- c_i * r_i^n mod n^2 with fresh r_i (phe/paillier.py:603-624): K1 with m = 0, then K3."""
        ctx = self.public_key.engine_context()
        count = len(self)
        if count:
            torch = _torch()
            d_r = torch.empty((count, ctx.n_limbs), dtype=torch.int32, device=self.limbs.device)
            ctx.random_lt_n_dev(d_r, count, stream=_stream(ctx))
            d_zero = torch.zeros((count, ctx.n_limbs), dtype=torch.int32, device=self.limbs.device)
            d_rn = torch.empty_like(self.limbs)
            ctx.encrypt_dev(d_zero, d_r, d_rn, count, stream=_stream(ctx))
            out = torch.empty_like(self.limbs)
            ctx.raw_add_dev(self.limbs, d_rn, out, count, stream=_stream(ctx))
            self.limbs = out
        self._obfuscated = True

    # ------------------------------------------------------------------ arithmetic
    def _raw_mul_rows(self, limbs, scalars):
        """limbs[i] ^ scalars[i] mod n^2 (device), scalars: list of ints in [0, n) or their limb matrix."""
        ctx = self.public_key.engine_context()
        torch = _torch()
        count = int(limbs.shape[0])
        d_s = _to_dev(scalars if isinstance(scalars, np.ndarray) else ints_to_limbs(scalars, ctx.n_limbs), ctx)
        out = torch.empty_like(limbs)
        status = torch.zeros((count,), dtype=torch.int32, device=limbs.device)
        ctx.raw_mul_dev(limbs, d_s, out, status, count, stream=_stream(ctx))
        if bool(status.any().item()):
            raise ZeroDivisionError('invert() no inverse exists')
        return out

    def decrease_exponent_to(self, new_exps):
        """Per-element exponent alignment: elements whose exponent is above new_exps[i] are raised to
        BASE^(delta) (phe/paillier.py:570-601); the others are untouched."""
        new_exps = np.broadcast_to(np.asarray(new_exps, dtype=np.int64), self.exponents.shape)
        if np.any(new_exps > self.exponents):
            raise ValueError('New exponent should be more negative than old exponent')
        idx = np.nonzero(new_exps < self.exponents)[0]
        limbs = self.limbs
        if len(idx):
            torch = _torch()
            ctx = self.public_key.engine_context()
            tidx = torch.from_numpy(idx).to(limbs.device)
            # the scalar BASE**delta is a small positive int: its encoding is the integer itself (exponent 0), so the
            # scalar limb matrix is built per distinct delta without touching EncodedNumber for every element
            deltas = (self.exponents[idx] - new_exps[idx]).astype(np.int64)
            scal = np.zeros((len(idx), ctx.n_limbs), dtype=np.uint32)
            for d in np.unique(deltas):
                factor = pow(EncodedNumber.BASE, int(d))
                if factor > self.public_key.max_int:
                    raise ValueError('Integer needs to be within +/- %d but got %d' % (self.public_key.max_int, factor))
                scal[deltas == d] = ints_to_limbs([factor], ctx.n_limbs)[0]
            sub = limbs[tidx].contiguous()
            out = torch.empty_like(sub)
            status = torch.zeros((len(idx),), dtype=torch.int32, device=limbs.device)
            ctx.raw_mul_dev(sub, _to_dev(scal, ctx), out, status, len(idx), stream=_stream(ctx))
            if bool(status.any().item()):
                raise ZeroDivisionError('invert() no inverse exists')
            limbs = limbs.clone()
            limbs[tidx] = out
        return EncryptedVector(self.public_key, limbs, new_exps.copy(), obfuscated=self._obfuscated and not len(idx))

    def __add__(self, other):
        ctx = self.public_key.engine_context()
        torch = _torch()
        if isinstance(other, EncryptedVector):
            if self.public_key != other.public_key:
                raise ValueError("Attempted to add numbers encrypted against different public keys!")
            if len(other) != len(self):
                raise ValueError("length mismatch")
            new_exps = np.minimum(self.exponents, other.exponents)
            a, b = self.decrease_exponent_to(new_exps), other.decrease_exponent_to(new_exps)
            out = torch.empty_like(a.limbs)
            if len(self):
                ctx.raw_add_dev(a.limbs, b.limbs, out, len(self), stream=_stream(ctx))
            return EncryptedVector(self.public_key, out, new_exps)
        # plaintext operand(s): encode against each element's exponent (phe/paillier.py:626-676)
        scalars = list(other) if hasattr(other, "__len__") else [other] * len(self)
        if len(scalars) != len(self):
            raise ValueError("length mismatch")
        n = self.public_key.n
        if any(isinstance(s, EncodedNumber) for s in scalars):
            encs = [s if isinstance(s, EncodedNumber) else EncodedNumber.encode(self.public_key, s, max_exponent=int(e))
                    for s, e in zip(scalars, self.exponents)]
            new_exps = np.minimum(self.exponents, np.array([e.exponent for e in encs], dtype=np.int64))
            a = self.decrease_exponent_to(new_exps)
            encs = [e.decrease_exponent_to(int(x)) if e.exponent > x else e for e, x in zip(encs, new_exps)]
            encodings = [e.encoding for e in encs]
        else:
            # plain numbers: one vectorised encode against every element's exponent; the encoded exponent never exceeds
            # it (max_exponent), so only `self` may need aligning
            s_limbs, new_exps = encode_batch(self.public_key, other if isinstance(other, np.ndarray) else scalars,
                                             max_exponent=self.exponents)
            a = self.decrease_exponent_to(new_exps)
            encodings = limbs_to_ints(s_limbs)
        nude = [(n * e + 1) % self.public_key.nsquare for e in encodings]          # raw_encrypt(., r=1), phe/paillier.py:673
        d_b = _to_dev(ints_to_limbs(nude, ctx.c_limbs), ctx)
        out = torch.empty_like(a.limbs)
        if len(self):
            ctx.raw_add_dev(a.limbs, d_b, out, len(self), stream=_stream(ctx))
        return EncryptedVector(self.public_key, out, new_exps)

    __radd__ = __add__

    def __mul__(self, other):
        if isinstance(other, EncryptedVector):
            raise NotImplementedError('Good luck with that...')
        scalars = list(other) if hasattr(other, "__len__") else [other] * len(self)
        if len(scalars) != len(self):
            raise ValueError("length mismatch")
        if any(isinstance(s, EncodedNumber) for s in scalars):
            encs = [s if isinstance(s, EncodedNumber) else EncodedNumber.encode(self.public_key, s) for s in scalars]
            s_limbs = ints_to_limbs([e.encoding for e in encs], self.public_key.engine_context().n_limbs)
            s_exps = np.array([e.exponent for e in encs], dtype=np.int64)
        else:
            s_limbs, s_exps = encode_batch(self.public_key, other if isinstance(other, np.ndarray) else scalars)
        out = self._raw_mul_rows(self.limbs, s_limbs) if len(self) else self.limbs
        return EncryptedVector(self.public_key, out, self.exponents + s_exps)

    __rmul__ = __mul__

    def __sub__(self, other):
        return self + (other * -1)

    def __rsub__(self, other):
        return other + (self * -1)

    def __truediv__(self, scalar):
        return self * (1 / scalar)

    def sum(self):
        """Homomorphic sum of all elements -> EncryptedNumber: the product of the ciphertexts modulo n^2 in two launches
        (pai_raw_sum: per-thread strided products, shared-memory tree per CTA, second launch over the CTA partials) --
        the reference's sum(list) / np.mean idiom (phe/tests/math_test.py:44-58) without B - 1 sequential _raw_add calls."""
        from .paillier import EncryptedNumber
        if not len(self):
            raise ValueError("empty vector")
        ctx = self.public_key.engine_context()
        torch = _torch()
        v = self.decrease_exponent_to(int(self.exponents.min()))
        out = torch.empty((1, v.limbs.shape[1]), dtype=torch.int32, device=v.limbs.device)
        ctx.raw_sum_dev(v.limbs, len(v), out, stream=_stream(ctx))
        return EncryptedNumber(self.public_key, limbs_to_ints(_to_host(out))[0], int(v.exponents[0]))

    def sum_chain(self):
        """The round-1 form of sum(): a chain of log2(B) pairwise raw_add launches (kept as the comparison baseline of
        bench.py's `reductions` leg and as a cross-check in the tests)."""
        from .paillier import EncryptedNumber
        if not len(self):
            raise ValueError("empty vector")
        ctx = self.public_key.engine_context()
        torch = _torch()
        v = self.decrease_exponent_to(int(self.exponents.min()))
        limbs = v.limbs
        while limbs.shape[0] > 1:
            half = limbs.shape[0] // 2
            out = torch.empty((half, limbs.shape[1]), dtype=torch.int32, device=limbs.device)
            ctx.raw_add_dev(limbs[:half].contiguous(), limbs[half:2 * half].contiguous(), out, half, stream=_stream(ctx))
            limbs = torch.cat([out, limbs[2 * half:]], dim=0) if limbs.shape[0] % 2 else out
        c = limbs_to_ints(_to_host(limbs))[0]
        return EncryptedNumber(self.public_key, c, int(v.exponents[0]))

    def _encode_scalars(self, scalars):
        if isinstance(scalars, EncryptedVector):
            raise NotImplementedError('Good luck with that...')
        sc = list(scalars) if not isinstance(scalars, np.ndarray) else scalars
        if len(sc) != len(self):
            raise ValueError("length mismatch")
        if not isinstance(sc, np.ndarray) and any(isinstance(x, EncodedNumber) for x in sc):
            encs = [x if isinstance(x, EncodedNumber) else EncodedNumber.encode(self.public_key, x) for x in sc]
            return (ints_to_limbs([e.encoding for e in encs], self.public_key.engine_context().n_limbs),
                    np.array([e.exponent for e in encs], dtype=np.int64))
        return encode_batch(self.public_key, sc)

    def dot(self, scalars):
        """Homomorphic dot product sum_i self[i] * scalars[i] -> EncryptedNumber (the encrypted scoring loop of
        examples/logistic_regression_encrypted_model.py:170-180): pai_raw_dot = Straus' simultaneous exponentiation over
        the group of elements each thread owns (one shared chain of squarings) + the product reduction of sum().
        Element exponents are aligned to the lowest one by folding BASE^delta into the plaintext scalar."""
        from .paillier import EncryptedNumber
        if not len(self):
            raise ValueError("empty vector")
        ctx = self.public_key.engine_context()
        torch = _torch()
        s_limbs, s_exps = self._encode_scalars(scalars)
        exps = self.exponents + s_exps
        emin = int(exps.min())
        up = np.nonzero(exps > emin)[0]
        if len(up):                                   # c^(k * BASE^delta) = (c^k)^(BASE^delta): alignment inside the exponent
            n, max_int = self.public_key.n, self.public_key.max_int
            vals = limbs_to_ints(s_limbs[up])
            new = []
            for k, d in zip(vals, (exps[up] - emin).tolist()):
                f = pow(EncodedNumber.BASE, int(d))
                mag = (n - k) if k >= n - max_int else k
                if mag * f > max_int:
                    raise ValueError('Integer needs to be within +/- %d but got %d' % (max_int, mag * f))
                new.append(k * f % n)
            s_limbs = s_limbs.copy()
            s_limbs[up] = ints_to_limbs(new, s_limbs.shape[1])
        out = torch.empty((1, self.limbs.shape[1]), dtype=torch.int32, device=self.limbs.device)
        status = torch.zeros((len(self),), dtype=torch.int32, device=self.limbs.device)
        ctx.raw_dot_dev(self.limbs, _to_dev(s_limbs, ctx), out, status, len(self), stream=_stream(ctx))
        if bool(status.any().item()):
            raise ZeroDivisionError('invert() no inverse exists')
        return EncryptedNumber(self.public_key, limbs_to_ints(_to_host(out))[0], emin)

    def dot_chain(self, scalars):
        """The round-1 form of dot(): one raw_mul launch, then sum_chain()."""
        return (self * scalars).sum_chain()

    # ------------------------------------------------------------------ wire format
    def to_json(self, be_secure=True):
        """The reference's basic JSON scheme (docs/serialisation.rst:24-31): {'public_key': {'n': ...},
        'values': [[str(ciphertext), exponent], ...]} -- readable by an unmodified phe peer."""
        if be_secure and not self._obfuscated:
            self.obfuscate()
        texts = limbs_to_decimal_strings(self.public_key.engine_context(), self.limbs)
        return '{"public_key": {"n": %d}, "values": [%s]}' % (
            self.public_key.n, ", ".join('["%s", %d]' % (t, e) for t, e in zip(texts, self.exponents.tolist())))

    @classmethod
    def from_json(cls, serialised, public_key=None):
        """Inverse of to_json (docs/serialisation.rst:35-42); ciphertexts go straight to the GPU."""
        import json
        from .paillier import PaillierPublicKey
        d = json.loads(serialised)
        pk = public_key or PaillierPublicKey(int(d["public_key"]["n"]))
        if pk.n != int(d["public_key"]["n"]):
            raise ValueError("serialised vector was encrypted against a different key")
        ctx = pk.engine_context()
        texts = [str(v[0]).lstrip("0") or "0" for v in d["values"]]
        bound = str(pk.nsquare)                    # the reference accepts any int; keep rows canonical (< n^2)
        texts = [t if len(t) < len(bound) or (len(t) == len(bound) and t < bound) or not t.isdigit()
                 else str(int(t) % pk.nsquare) for t in texts]
        d_c = decimal_strings_to_limbs(ctx, texts, ctx.c_limbs)
        return cls(pk, d_c, [int(v[1]) for v in d["values"]])

    # ------------------------------------------------------------------ decryption
    def decrypt_encoded(self, private_key):
        if self.public_key != private_key.public_key:
            raise ValueError('encrypted_number was encrypted against a different key!')
        ctx = private_key.engine_context()
        torch = _torch()
        count = len(self)
        plain = limbs_to_ints(self._decrypt_rows(ctx)) if count else []
        return [EncodedNumber(self.public_key, m, int(e)) for m, e in zip(plain, self.exponents)]

    def decrypt(self, private_key):
        """Decrypt and decode (vectorised EncodedNumber.decode) -> list of Python floats / ints."""
        if self.public_key != private_key.public_key:
            raise ValueError('encrypted_number was encrypted against a different key!')
        ctx = private_key.engine_context()
        torch = _torch()
        count = len(self)
        if not count:
            return []
        return decode_batch(self.public_key, self._decrypt_rows(ctx), self.exponents)

    def _decrypt_rows(self, ctx):
        """raw_decrypt of every row -> host uint32 matrix [B, n_limbs of the PUBLIC layout].  The private context sizes
        its rows from max(p, q), the public one from n: for unbalanced primes they differ, and the rows are re-padded."""
        torch = _torch()
        count = len(self)
        d_c = _rows_for(ctx.c_limbs, self.limbs)
        d_m = torch.empty((count, ctx.n_limbs), dtype=torch.int32, device=self.limbs.device)
        ctx.decrypt_dev(d_c, d_m, count, stream=_stream(ctx))
        return _to_host(_rows_for(self.public_key.engine_context().n_limbs, d_m))
