"""Fixed-point encoding of ints/floats as integers mod n -- same behaviour as phe/encoding.py.

An ``EncodedNumber`` stores ``encoding`` in [0, n) and an ``exponent`` so that the value is
``mantissa * BASE**exponent`` where the mantissa is ``encoding`` (positive, <= max_int) or
``encoding - n`` (negative, >= n - max_int); the middle third of [0, n) signals overflow
(phe/encoding.py:110-233).  This is cheap host arithmetic on either side of the kernels; the
batched path vectorises it in ``vector.py``.
"""
import fractions
import math
import sys


class EncodedNumber(object):
    BASE = 16
    LOG2_BASE = math.log(BASE, 2)
    FLOAT_MANTISSA_BITS = sys.float_info.mant_dig

    def __init__(self, public_key, encoding, exponent):
        self.public_key = public_key
        self.encoding = encoding
        self.exponent = exponent

    @classmethod
    def _precision_exponent(cls, scalar, precision):
        """Largest exponent that still represents `scalar` (or `precision`) exactly enough
        (phe/encoding.py:160-176)."""
        if precision is not None:
            return math.floor(math.log(precision, cls.BASE))
        if isinstance(scalar, int):
            return 0
        if isinstance(scalar, float):
            lsb_exponent = math.frexp(scalar)[1] - cls.FLOAT_MANTISSA_BITS      # weight of the last mantissa bit
            return math.floor(lsb_exponent / cls.LOG2_BASE)
        raise TypeError("Don't know the precision of type %s." % type(scalar))

    @classmethod
    def encode(cls, public_key, scalar, precision=None, max_exponent=None):
        exponent = cls._precision_exponent(scalar, precision)
        if max_exponent is not None:
            exponent = min(max_exponent, exponent)
        scaled = fractions.Fraction(scalar) * fractions.Fraction(cls.BASE) ** -exponent   # exact rationals, no float overflow
        int_rep = round(scaled)
        if abs(int_rep) > public_key.max_int:
            raise ValueError('Integer needs to be within +/- %d but got %d' % (public_key.max_int, int_rep))
        return cls(public_key, int_rep % public_key.n, exponent)

    def decode(self):
        n, max_int, enc = self.public_key.n, self.public_key.max_int, self.encoding
        if enc >= n:
            raise ValueError('Attempted to decode corrupted number')
        if enc <= max_int:
            mantissa = enc
        elif enc >= n - max_int:
            mantissa = enc - n
        else:
            raise OverflowError('Overflow detected in decrypted number')
        if self.exponent >= 0:
            return mantissa * self.BASE ** self.exponent
        try:
            return mantissa / self.BASE ** -self.exponent
        except OverflowError as e:
            raise OverflowError('decoded result too large for a float') from e

    def decrease_exponent_to(self, new_exp):
        if new_exp > self.exponent:
            raise ValueError('New exponent %i should be more negative than'
                             'old exponent %i' % (new_exp, self.exponent))
        factor = pow(self.BASE, self.exponent - new_exp)
        return self.__class__(self.public_key, self.encoding * factor % self.public_key.n, new_exp)
