#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Pulls the two complete C programs (and the sample MPS file) out of the reference's C-API
documentation, docs/cuopt/source/cuopt-c/lp-milp/lp-example.rst, into a scratch directory so that oracle/Makefile can
compile them UNCHANGED against this repo's headers and library (tests/test_reference_doc_examples.py runs them).
Nothing is copied into the repository: the outputs live under the git-ignored oracle/_ref/.
usage: extract_doc_examples.py <lp-example.rst> <out_dir>"""
import os
import sys


def blocks(lines, marker):
    """Indented bodies of the `.. code-block:: <marker>` directives, in order."""
    out, i = [], 0
    while i < len(lines):
        if lines[i].strip() == f".. code-block:: {marker}":
            i += 1
            while i < len(lines) and (not lines[i].strip() or lines[i].lstrip().startswith(":")):
                i += 1
            body = []
            while i < len(lines) and (not lines[i].strip() or lines[i].startswith("   ")):
                body.append(lines[i][3:] if lines[i].startswith("   ") else "")
                i += 1
            out.append("\n".join(body).rstrip() + "\n")
        else:
            i += 1
    return out


def main():
    rst, out_dir = sys.argv[1], sys.argv[2]
    os.makedirs(out_dir, exist_ok=True)
    with open(rst) as f:
        lines = f.read().split("\n")
    c_blocks = blocks(lines, "c")
    assert len(c_blocks) >= 2, "expected the two example programs"
    for name, src in zip(("doc_lp_example.c", "doc_lp_example_mps.c"), c_blocks):
        with open(os.path.join(out_dir, name), "w") as f:
            f.write(src)
    # the MPS file of the second example: between  echo "  and  " > sample.mps  of a bash block
    for b in blocks(lines, "bash"):
        if "> sample.mps" in b and 'echo "' in b:
            body = b[b.index('echo "') + 6: b.index('" > sample.mps')]
            with open(os.path.join(out_dir, "doc_sample.mps"), "w") as f:
                f.write(body + "\n")
            break
    else:
        raise SystemExit("sample.mps block not found")


if __name__ == "__main__":
    main()
