"""ctypes calls carry no signature: a Python call site with one argument too few would silently pass garbage to a kernel launcher.
This test parses the `extern "C"` prototypes of the native sources and the call sites in the package and compares arities."""

import pathlib
import re

ROOT = pathlib.Path(__file__).resolve().parent.parent / "aggregathor_b200"


def _prototypes():
  found = {}
  for path in list((ROOT / "native").rglob("*.cu")) + list((ROOT / "native").rglob("*.cpp")):
    source = path.read_text()
    for match in re.finditer(r"\n(?:int|long long|char const\*)\s+(agb_\w+)\s*\(([^)]*)\)\s*\{", source):
      found[match.group(1)] = len([a for a in match.group(2).split(",") if a.strip()])
    for match in re.finditer(r"extern \"C\" int (agb_\w+)\(([^)]*)\)", source):   # macro-generated exports (host GARs)
      found.setdefault(match.group(1), len([a for a in match.group(2).split(",") if a.strip()]))
  return found


def _call_arity(source, start):
  """Number of top-level arguments of the call whose '(' is at `start`; '*name' counts as the length of the tuple literal `name = (...)`."""
  depth, pieces, current, i = 0, [], "", start
  while i < len(source):
    ch = source[i]
    if ch in "([{":
      depth += 1
      if depth > 1:
        current += ch
    elif ch in ")]}":
      depth -= 1
      if depth == 0:
        break
      current += ch
    elif ch == "," and depth == 1:
      pieces.append(current)
      current = ""
    else:
      current += ch
    i += 1
  if current.strip():
    pieces.append(current)
  total = 0
  for piece in pieces:
    piece = piece.strip()
    if piece.startswith("*") and re.fullmatch(r"\*\w+", piece):
      literal = re.search(piece[1:] + r" = \(", source)
      assert literal is not None, piece
      total += _call_arity(source, literal.end() - 1)
    else:
      total += 1
  return total


def test_every_native_call_matches_its_prototype():
  prototypes = _prototypes()
  assert len(prototypes) >= 45
  checked = 0
  for path in ROOT.rglob("*.py"):
    source = path.read_text()
    for match in re.finditer(r"\.(agb_\w+)\(", source):
      name = match.group(1)
      if name.endswith(("_float", "_double")) or name not in prototypes:
        continue
      assert _call_arity(source, match.end() - 1) == prototypes[name], (path.name, name)
      checked += 1
  assert checked >= 30
