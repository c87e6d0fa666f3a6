#!/usr/bin/env python
"""Build a Python-3-importable scratch copy of the reference's model-build code.

The reference (/root/reference/src/IMCoalHMM) is Python 2.  This tool copies the
package into a scratch directory OUTSIDE the repository (default
/tmp/imcoalhmm_ref_shim) and applies six purely mechanical py2->py3 edits
 so the reference's own `build_hidden_markov_model` can be
executed here to produce golden (pi, T, E) vectors.  No reference source is ever
written into this repository; only the numbers it produces are committed (by
tools/gen_golden.py, under tests/golden/).

The edits:
  1. xrange -> range
  2. .iteritems() -> .items()
  3. `from cache import Cache` -> `from IMCoalHMM.cache import Cache`
  4. `from scipy import matrix` -> `from numpy import matrix`
  5. `x = map(ComputeThroughInterval(...), ...)` -> `x = list(map(...))`
  6. py2 print statements -> print() calls (only in demo `main()`s / log callback)
  7. (with_hmm=True only) hmm.py: `np.array(map(int, ...))` -> `np.array(list(map(int, ...)))`; hmm.py is the 22-line
     wrapper around the absent `ziphmm` module (hmm.py:7,16,20-21) -- importing it needs a `ziphmm` in sys.modules,
     which is exactly how tests/test_reference_boundary.py runs the reference's own code against this package
"""
import os
import re
import shutil
import sys

REFERENCE_PKG = "/root/reference/src/IMCoalHMM"
DEFAULT_OUT = "/tmp/imcoalhmm_ref_shim"

# Modules that need the absent ziphmm / pyZipHMM extension or run demos at import
# time; they are not part of the model-build path and are left out of the shim.
SKIP = {"hmm.py", "mcmc.py", "ILS.py", "admixture.py", "genetic_algorithm.py", "particle_swarm.py"}

_PRINT_STMT = re.compile(r"^(\s*)print\b(?!\s*\()(.*)$")


def _fix_print(line):
    m = _PRINT_STMT.match(line)
    if not m:
        return line
    indent, rest = m.group(1), m.group(2).strip()
    trailing_comma = rest.endswith(",")
    if trailing_comma:
        rest = rest[:-1].rstrip()
    if rest.startswith(">>"):
        target, _, payload = rest[2:].partition(",")
        args = payload.strip()
        extra = "file=%s" % target.strip()
        body = (args + ", " + extra) if args else extra
    else:
        body = rest
    if trailing_comma:
        body = (body + ", end=' '") if body else "end=' '"
    return "%sprint(%s)\n" % (indent, body)


def convert(text):
    text = text.replace("xrange", "range")
    text = text.replace(".iteritems()", ".items()")
    text = text.replace("from cache import Cache", "from IMCoalHMM.cache import Cache")
    text = text.replace("from scipy import matrix", "from numpy import matrix")
    text = re.sub(r"= map\((ComputeThroughInterval\(.*?\),\s*range\([^\n]*?\))\)\n",
                  r"= list(map(\1))\n", text, flags=re.S)
    text = text.replace("np.array(map(int, finp.read().split()), dtype=np.int32)",
                        "np.array(list(map(int, finp.read().split())), dtype=np.int32)")
    return "".join(_fix_print(l) for l in text.splitlines(True))


def build(out_dir=DEFAULT_OUT, reference_pkg=REFERENCE_PKG, with_hmm=False):
    pkg_out = os.path.join(out_dir, "IMCoalHMM")
    if os.path.isdir(out_dir):
        shutil.rmtree(out_dir)
    os.makedirs(pkg_out)
    for name in sorted(os.listdir(reference_pkg)):
        if not name.endswith(".py") or (name in SKIP and not (with_hmm and name == "hmm.py")):
            continue
        with open(os.path.join(reference_pkg, name)) as f:
            src = f.read()
        with open(os.path.join(pkg_out, name), "w") as f:
            f.write(convert(src))
    return out_dir


def import_shim(out_dir=DEFAULT_OUT):
    """Build (if needed) and put the shim at the front of sys.path."""
    if not os.path.isdir(os.path.join(out_dir, "IMCoalHMM")):
        build(out_dir)
    if out_dir not in sys.path:
        sys.path.insert(0, out_dir)


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else DEFAULT_OUT
    print("reference shim written to", build(out))
