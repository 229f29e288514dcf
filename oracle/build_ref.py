#!/usr/bin/env python3
"""
ORACLE / checker infrastructure (never imported by gpt_b200/): compiles the reference's own scripts for the hot path, from the
sources where they lie under /root/reference, into Python code objects under oracle/_ref/ (git-ignored binaries, like a .so;
they travel to the GPU box where /root/reference does not exist).  tests/test_reference_scripts.py executes them UNCHANGED over
this repo's `gpt` / `cgpt` modules:

  benchmarks/dslash.py, benchmarks/wilson_clover_dslash.py     whole files
  tests/qcd/fermion_operators.py                               three blocks of top-level statements, untouched: the set-up (lines up
      to `grid = U[0].grid`), everything from `wilson_params = ...` up to the test-suite loop (parameter tables with the golden
      finger prints, the verify_* helpers), and the loop itself.  The blocks in between (the pure-Python reference Wilson operator,
      propagator / split-grid / staggered tests) exercise parts of GPT outside the hot path.
  lib/gpt/qcd/fermion/operator/interface.py, lib/gpt/qcd/fermion/register.py, lib/gpt/algorithms/inverter/cg.py
                                                               whole files: the reference's Python glue closest to the boundary
      (handle cache + opcode dispatch over `cgpt`, and the CG loop), run over this repo's cgpt stand-in and core objects

No source text is copied: only marshalled code objects are written.  Run by __graft_entry__.build() when the reference tree is
present; a no-op otherwise.
"""
import ast
import marshal
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("GPT_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")

WHOLE = [
    "benchmarks/dslash.py", "benchmarks/wilson_clover_dslash.py",
    "lib/gpt/qcd/fermion/operator/interface.py", "lib/gpt/qcd/fermion/register.py",
    "lib/gpt/algorithms/inverter/cg.py",
]


def _dump(code, name):
    with open(os.path.join(OUT, name + ".code"), "wb") as f:
        marshal.dump(code, f)


def _assigns(node, name):
    return isinstance(node, ast.Assign) and any(isinstance(t, ast.Name) and t.id == name for t in node.targets)


def build():
    if not os.path.isdir(REFERENCE):
        return False
    os.makedirs(OUT, exist_ok=True)
    for rel in WHOLE:
        path = os.path.join(REFERENCE, rel)
        with open(path) as f:
            _dump(compile(f.read(), path, "exec"), rel.replace("/", "__").replace(".py", ""))
    # blocks of tests/qcd/fermion_operators.py, selected by what the statements are, not by line numbers
    path = os.path.join(REFERENCE, "tests/qcd/fermion_operators.py")
    with open(path) as f:
        body = ast.parse(f.read(), path).body
    i_head = next(i for i, n in enumerate(body) if _assigns(n, "grid"))  # `grid = U[0].grid`
    i_defs = next(i for i, n in enumerate(body) if _assigns(n, "wilson_params"))
    i_loop = next(i for i, n in enumerate(body) if isinstance(n, ast.For) and isinstance(n.iter, ast.Name) and n.iter.id == "test_suite")
    blocks = {"head": body[:i_head + 1], "defs": body[i_defs:i_loop], "loop": body[i_loop:i_loop + 1]}
    for tag, nodes in blocks.items():
        mod = ast.Module(body=nodes, type_ignores=[])
        _dump(compile(mod, path, "exec"), "tests__qcd__fermion_operators__" + tag)
    with open(os.path.join(OUT, "MANIFEST"), "w") as f:
        f.write(f"python {sys.version_info.major}.{sys.version_info.minor}; compiled from {REFERENCE} by oracle/build_ref.py\n")
        for n in sorted(os.listdir(OUT)):
            if n.endswith(".code"):
                f.write(n + "\n")
    return True


def load(name):
    """code object by name (e.g. "benchmarks__dslash"), or None if the checker has not been built"""
    path = os.path.join(OUT, name + ".code")
    if not os.path.exists(path):
        return None
    with open(path, "rb") as f:
        return marshal.load(f)


if __name__ == "__main__":
    print("built" if build() else f"{REFERENCE} not present: nothing to do")
