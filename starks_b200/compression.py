"""Proof (de)serialisation with the wire format of starks/compression.py (the only
serialisation in the reference): a proof is flattened to a list of byte
strings, framed by the 4-byte markers b'----' b'++++' b'====' b'////', and every repeated
object is replaced by a 2-byte big-endian index of its first occurrence.

Own design, same bytes: a proof is first turned into a flat TOKEN STREAM (generators below),
one pass de-duplicates the stream, and decoding is the two inverse passes -- resolve every
back-reference, then split the resolved stream on the markers -- instead of the reference's
positional cursor loops.  Host-side, O(objects)."""

LAYER, BRANCH_END, QUERY_END, FINAL = b"----", b"++++", b"====", b"////"


def _dedup(tokens):
  """First occurrence verbatim, later ones as 2-byte positions (compression.py:9-14).  Raises
  OverflowError past 65 535 objects, like the reference's to_bytes(2, 'big')."""
  out, first = [], {}
  for tok in tokens:
    at = first.setdefault(tok, len(out))
    out.append(tok if at == len(out) else at.to_bytes(2, "big"))
  return out


def _resolve(stream):
  """Inverse of _dedup: any 2-byte item is a position in the stream (compression.py:36-38)."""
  return [stream[int.from_bytes(tok, "big")] if len(tok) == 2 else tok for tok in stream]


def _split(tokens, marker):
  """Runs of tokens terminated by `marker`; a trailing unterminated run is kept (the
  reference's cursor loops would return it too), an empty tail is not."""
  runs, cur = [], []
  for tok in tokens:
    if tok == marker:
      runs.append(cur)
      cur = []
    else:
      cur.append(tok)
  if cur:
    runs.append(cur)
  return runs


def _fri_tokens(prf):
  for root, queries in prf[:-1]:
    yield LAYER
    yield root
    for query in queries:
      for branch in query:
        yield from branch
        yield BRANCH_END
      yield QUERY_END
  yield FINAL
  yield from prf[-1]


def _branch_tokens(branches):
  for branch in branches:
    yield from branch
    yield LAYER


def compress_fri(prf):
  """compression.py:1-31: [[root, [[branch, ...], ...]], ..., final values] -> flat list."""
  out = _dedup(_fri_tokens(prf))
  assert decompress_fri(out) == prf
  return out


def decompress_fri(proof):
  """compression.py:34-64."""
  cut = proof.index(FINAL)          # the terminator occurs once, so it is never a back-reference
  toks = _resolve(proof)
  layers = []
  for run in _split(toks[:cut] + [LAYER], LAYER)[1:]:      # run = root, then the queries
    queries = [_split(q, BRANCH_END) for q in _split(run[1:], QUERY_END)]
    layers.append([run[0], queries])
  layers.append(toks[cut + 1:])
  return layers


def compress_branches(branches):
  """compression.py:67-85."""
  out = _dedup(_branch_tokens(branches))
  assert decompress_branches(out) == branches
  return out


def decompress_branches(proof):
  """compression.py:88-103."""
  toks = _resolve(proof)
  runs = _split(toks, LAYER)
  # an empty branch between two markers is a run too; _split keeps those, and drops only the
  # empty tail after the last marker
  return runs


def bin_length(c):
  """compression.py:106-107: serialised size, 32-byte objects carry a one-byte 0xff tag."""
  return sum(len(x) + (len(x) == 32) for x in c)
