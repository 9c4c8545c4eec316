"""Host-side value type for Z/p, mirroring starks/modp.py:25-106 (IntegersModP class factory:
canonical residue in .n, 32-byte big-endian to_bytes, bytes constructor WITHOUT reduction).
Only protocol glue and API conversions use it; all bulk arithmetic runs on the device."""
import functools


@functools.lru_cache(maxsize=None)
def IntegersModP(p):
  class IntegerModP(object):
    __slots__ = ("n",)

    def __init__(self, n):
      if isinstance(n, bytes):
        self.n = int.from_bytes(n, "big")            # modp.py:33-34: not reduced
      elif isinstance(n, IntegerModP):
        self.n = n.n
      else:
        try:
          self.n = int(n) % IntegerModP.p            # modp.py:35-36
        except Exception:
          raise TypeError("Can't cast type %s to %s in __init__" % (type(n).__name__, type(self).__name__))

    @property
    def field(self):
      return IntegerModP

    @staticmethod
    def _coerce(other):
      if isinstance(other, IntegerModP):
        return other
      if isinstance(other, int):
        return IntegerModP(other)
      return None

    def __add__(self, other):
      o = self._coerce(other)
      return NotImplemented if o is None else IntegerModP(self.n + o.n)
    __radd__ = __add__

    def __sub__(self, other):
      o = self._coerce(other)
      return NotImplemented if o is None else IntegerModP(self.n - o.n)

    def __rsub__(self, other):
      o = self._coerce(other)
      return NotImplemented if o is None else IntegerModP(o.n - self.n)

    def __mul__(self, other):
      o = self._coerce(other)
      return NotImplemented if o is None else IntegerModP(self.n * o.n)
    __rmul__ = __mul__

    def __neg__(self):
      return IntegerModP(-self.n)

    def __eq__(self, other):
      o = self._coerce(other)
      return o is not None and self.n == o.n

    def __ne__(self, other):
      return not self.__eq__(other)

    def __hash__(self):
      return hash((self.n, IntegerModP.p))

    def inverse(self):
      return IntegerModP(pow(self.n, -1, IntegerModP.p))  # modp.py:71-79 (ext. Euclid)

    def __truediv__(self, other):
      o = self._coerce(other)
      return NotImplemented if o is None else self * o.inverse()

    def __rtruediv__(self, other):
      o = self._coerce(other)
      return NotImplemented if o is None else o * self.inverse()

    def __pow__(self, e):
      return IntegerModP(pow(self.n, int(e), IntegerModP.p))  # numbertype.py:68-84

    def __abs__(self):
      return abs(self.n)

    def __int__(self):
      return self.n

    def __index__(self):
      return self.n

    def __str__(self):
      return str(self.n)

    def __repr__(self):
      return "%d (mod %d)" % (self.n, IntegerModP.p)

    def to_bytes(self):
      return self.n.to_bytes(32, "big")            # modp.py:94-95

  IntegerModP.p = p
  IntegerModP.m = 1
  IntegerModP.field_size = p
  IntegerModP.__name__ = "Z/%d" % p
  return IntegerModP


def element_to_int(x):
  """int / IntegerModP-like (has .n) / 32-byte big-endian bytes -> int."""
  if isinstance(x, int):
    return x
  n = getattr(x, "n", None)
  if n is not None:
    return int(n)
  if isinstance(x, (bytes, bytearray)):
    return int.from_bytes(x, "big")
  return int(x)
