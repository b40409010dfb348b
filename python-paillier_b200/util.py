"""The scalar bigint seam of the reference, routed to the CUDA engine.

Mirrors the three functions ``phe.paillier`` imports by name (phe/paillier.py:29): ``powmod``
(phe/util.py:38-50), ``mulmod`` (:53-64), ``invert`` (:85-103), with the same argument meaning,
return values (canonical residues as Python ints) and errors (``ZeroDivisionError``).  Where the
reference switches to gmpy2 for large operands (``_USE_MOD_FROM_GMP_SIZE``, :35-36, 47, 60), this
module switches to the GPU engine (batch of one); small operands and even moduli -- neither occurs
on the Paillier path -- use Python ints exactly like the reference's own small-operand branch.
Key generation helpers (out of the hot-path scope, phe/util.py:106-161, 381-443) are plain Python.
"""
import math
import random
import threading
from base64 import urlsafe_b64decode, urlsafe_b64encode
from collections import OrderedDict

from . import engine as _engine

# operands below this size stay on Python ints, as in the reference (phe/util.py:35-36)
_USE_MOD_FROM_GMP_SIZE = (1 << (8 * 2))
_USE_MULMOD_FROM_GMP_SIZE = (1 << 1000)
_MAX_ENGINE_BITS = 8192

_ctx_cache = OrderedDict()
_ctx_lock = threading.Lock()
_CTX_CACHE_SIZE = 16


def _mod_ctx(modulus):
    """Least-recently-used cache of Montgomery contexts.  An evicted context is only dropped from the cache: another
    thread may still be inside a call on it, so it is destroyed by its finaliser once the last reference is gone."""
    with _ctx_lock:
        ctx = _ctx_cache.get(modulus)
        if ctx is None:
            ctx = _engine.ModContext(modulus)
            _ctx_cache[modulus] = ctx
            while len(_ctx_cache) > _CTX_CACHE_SIZE:
                _ctx_cache.popitem(last=False)
        else:
            _ctx_cache.move_to_end(modulus)
        return ctx


def _engine_modulus(c):
    return c > _USE_MOD_FROM_GMP_SIZE and (c & 1) == 1 and c.bit_length() <= _MAX_ENGINE_BITS


def powmod(a, b, c):
    """a ** b mod c."""
    if a == 1:                                           # phe/util.py:45-46
        return 1
    if b < 0 or not _engine_modulus(c) or max(a, b, c) < _USE_MOD_FROM_GMP_SIZE:
        return pow(a, b, c)
    ctx = _mod_ctx(c)
    if a < 0 or a.bit_length() > 64 * ctx.limbs:
        a %= c
    return ctx.powmod([a], b)[0]


def mulmod(a, b, c):
    """a * b mod c, non-negative also for negative a (crt passes mq - mp, phe/paillier.py:373)."""
    if not _engine_modulus(c) or max(a, b, c) < _USE_MULMOD_FROM_GMP_SIZE:
        return a * b % c
    ctx = _mod_ctx(c)
    lim = 1 << (32 * ctx.limbs)
    if not 0 <= a < lim:
        a %= c
    if not 0 <= b < lim:
        b %= c
    return ctx.mulmod([a], [b])[0]


def extended_euclidean_algorithm(a, b):
    """(r, s, t) with r = gcd(a, b) = s*a + t*b."""
    r_prev, r = a, b
    s_prev, s = 1, 0
    t_prev, t = 0, 1
    while r:
        k = r_prev // r
        r_prev, r = r, r_prev - k * r
        s_prev, s = s, s_prev - k * s
        t_prev, t = t, t_prev - k * t
    return r_prev, s_prev, t_prev


def invert(a, b):
    """Multiplicative inverse of a modulo b; ZeroDivisionError if there is none."""
    if _engine_modulus(b):
        ctx = _mod_ctx(b)
        if not 0 <= a < (1 << (32 * ctx.limbs)):
            a %= b
        out, status = ctx.invert([a])
        if status[0]:
            raise ZeroDivisionError('invert() no inverse exists')
        return out[0]
    g, s, _ = extended_euclidean_algorithm(a, b)
    if g != 1:
        raise ZeroDivisionError('invert() no inverse exists')
    return s % b


# ----------------------------------------------------------------------------- key generation (not hot path)
_SMALL_PRIMES = [p for p in range(2, 2000) if all(p % d for d in range(2, int(p ** 0.5) + 1))]
first_primes = _SMALL_PRIMES


def miller_rabin(n, k):
    """k rounds of Miller-Rabin with random bases; False = composite, True = probably prime."""
    assert n > 3
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    rnd = random.SystemRandom()
    for _ in range(k):
        x = pow(rnd.randrange(2, n - 1), d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def is_prime(n, mr_rounds=25):
    if n <= _SMALL_PRIMES[-1]:
        return n in _SMALL_PRIMES
    if any(n % p == 0 for p in _SMALL_PRIMES):
        return False
    return miller_rabin(n, mr_rounds)


def is_prime_batch(candidates, mr_rounds=25):
    """is_prime (phe/util.py:420-443) for a list of candidates: trial division by the small primes on the host, then ALL
    survivors through one batched Miller-Rabin launch on the device (engine.miller_rabin_batch)."""
    out = [None] * len(candidates)
    todo = []
    for i, n in enumerate(candidates):
        if n <= _SMALL_PRIMES[-1]:
            out[i] = n in _SMALL_PRIMES
        elif any(n % p == 0 for p in _SMALL_PRIMES):
            out[i] = False
        else:
            todo.append(i)
    if todo:
        for i, r in zip(todo, _engine.miller_rabin_batch([candidates[i] for i in todo], mr_rounds)):
            out[i] = r
    return out


def getprimeover_batch(N, count=1, width=None):
    """`count` random N-bit primes: windows of random odd candidates are sieved on the host and tested together on the
    device until enough primes are found (the batched form of getprimeover, phe/util.py:106-124)."""
    rnd = random.SystemRandom()
    primes = []
    width = width or max(64, 24 * count + N // 8)
    while len(primes) < count:
        cands = [rnd.randrange(1 << (N - 1), 1 << N) | 1 for _ in range(width)]
        primes += [c for c, ok in zip(cands, is_prime_batch(cands)) if ok]
    return primes[:count]


def getprimeover(N):
    """A random N-bit prime from the system's CSPRNG, one candidate at a time on the host like the reference
    (phe/util.py:106-124); many primes at once: getprimeover_batch (device)."""
    rnd = random.SystemRandom()
    cand = rnd.randrange(1 << (N - 1), 1 << N) | 1
    while not is_prime(cand):
        cand += 2
    return cand


def isqrt(N):
    return math.isqrt(N)


def improved_i_sqrt(n):
    assert n >= 0
    return math.isqrt(n)


# ----------------------------------------------------------------------------- serialisation helpers
def base64url_encode(payload):
    if not isinstance(payload, bytes):
        payload = payload.encode('utf-8')
    return urlsafe_b64encode(payload).decode('utf-8').rstrip('=')


def base64url_decode(payload):
    payload += '=' * (-len(payload) % 4)
    return urlsafe_b64decode(payload.encode('utf-8'))


def base64_to_int(source):
    return int.from_bytes(base64url_decode(source), 'big')


def int_to_base64(source):
    assert source != 0
    return base64url_encode(source.to_bytes((source.bit_length() + 7) // 8, 'big'))
