"""Mirror of starks/compression.py (the only serialisation in the reference): proofs are
flattened to a list of byte strings in which a repeated object is replaced by a 2-byte
big-endian back-reference to its first position (compression.py:9-14), with the framing
markers b'----', b'++++', b'====', b'////'.  Host-side; SURVEY.md 8(f) rank 1."""


class _Dedup(object):
  """compression.py:6-14 / 70-78: first occurrence verbatim, later ones as 2-byte indices."""

  def __init__(self):
    self.out, self.index = [], {}

  def add(self, x):
    if x in self.index:
      self.out.append(self.index[x].to_bytes(2, "big"))   # OverflowError past 65535 objects, as upstream
    else:
      self.out.append(x)
      self.index[x] = len(self.out) - 1


def _deref(proof, pos):
  item = proof[pos]
  return proof[int.from_bytes(item, "big")] if len(item) == 2 else item


def compress_fri(prf):
  """compression.py:1-31."""
  d = _Dedup()
  for root, yproofs in prf[:-1]:
    d.add(b"----")
    d.add(root)
    for yproof in yproofs:
      for branch in yproof:
        for p in branch:
          d.add(p)
        d.add(b"++++")
      d.add(b"====")
  d.add(b"////")
  for x in prf[-1]:
    d.add(x)
  assert decompress_fri(d.out) == prf
  return d.out


def decompress_fri(proof):
  """compression.py:34-64."""
  o, pos = [], 0
  while proof[pos] != b"////":
    assert _deref(proof, pos) == b"----"
    root = _deref(proof, pos + 1)
    pos += 2
    yproofs = []
    while _deref(proof, pos) not in (b"----", b"////"):
      yproof = []
      while _deref(proof, pos) != b"====":
        branch = []
        while _deref(proof, pos) != b"++++":
          branch.append(_deref(proof, pos))
          pos += 1
        yproof.append(branch)
        pos += 1
      yproofs.append(yproof)
      pos += 1
    o.append([root, yproofs])
  pos += 1
  o.append([_deref(proof, x) for x in range(pos, len(proof))])
  return o


def compress_branches(branches):
  """compression.py:67-85."""
  d = _Dedup()
  for branch in branches:
    for p in branch:
      d.add(p)
    d.add(b"----")
  assert decompress_branches(d.out) == branches
  return d.out


def decompress_branches(proof):
  """compression.py:88-103."""
  o, pos = [], 0
  while pos < len(proof):
    branch = []
    while pos < len(proof) and _deref(proof, pos) != b"----":
      branch.append(_deref(proof, pos))
      pos += 1
    o.append(branch)
    pos += 1
  return o


def bin_length(c):
  """compression.py:106-107."""
  return len(b"".join([(b"\xff" if len(x) == 32 else b"") + x for x in c]))
