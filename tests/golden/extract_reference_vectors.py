#!/usr/bin/env python3
"""
Extract the golden vectors / known-answer values the reference's own tests hold for the fermion-operator hot path
 from /root/reference and write them to tests/golden/reference_vectors.json.

  python tests/golden/extract_reference_vectors.py [/root/reference]

The reference cannot be executed in this environment (its arithmetic lives in Grid, which is neither vendored nor
installable offline), so the fixtures are the numbers its tests assert, copied by this script together with the file and
line they come from.  tests/test_oracle_golden.py::test_fixture_matches_embedded_numbers keeps the numbers used in the
parity tests identical to this file; where /root/reference is present it also re-runs the extraction.
"""
import ast
import json
import os
import re
import sys


def _literal(node):
    """python literal of an ast node, with 1j arithmetic evaluated (the reference writes a + b * 1j)"""
    return eval(compile(ast.Expression(node), "<golden>", "eval"), {"__builtins__": {}})


def _assignments(path):
    src = open(path).read()
    tree = ast.parse(src)
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
            try:
                out[node.targets[0].id] = (_literal(node.value), node.lineno)
            except Exception:
                pass
    return out


def _c(z):
    z = complex(z)
    return [z.real, z.imag]


def extract(root):
    vec = {}
    # ---- operator fingerprints <dst| M |src> --------------------------------------------------------------------------
    f = os.path.join(root, "tests/qcd/fermion_operators.py")
    a = _assignments(f)
    fp = {}
    for name in ["wilson_matrices", "wilson_clover_matrices", "wilson_clover_matrices_open", "wilson_twisted_mass_matrices"]:
        val, line = a[name]
        fp[name] = {"values": {k: _c(v[0]) for k, v in val.items()}, "source": f"tests/qcd/fermion_operators.py:{line}"}
    for name in ["wilson_params", "wilson_clover_params", "wilson_clover_open_params", "wilson_twisted_mass_params"]:
        if name in a:
            fp[name] = {"values": a[name][0], "source": f"tests/qcd/fermion_operators.py:{a[name][1]}"}
    # the test_suite dict holds callables (g.qcd.fermion.*): read the "matrices" sub-dicts with a regular expression
    src = open(f).read()
    for m in re.finditer(r'"(zmobius|mobius|mobius_axial_mass)"\s*:\s*\{', src):
        name = m.group(1)
        sub = src[m.end():]
        mm = re.search(r'"matrices"\s*:\s*\{(.*?)\n\s*\},', sub, re.S)
        vals = {}
        for key, re_, im_ in re.findall(r'"([^"]*)"\s*:\s*\[\s*([-0-9.e]+)\s*([-+]\s*[0-9.e]+)j\s*\]', mm.group(1)):
            vals[key] = [float(re_), float(im_.replace(" ", ""))]
        line = src[:m.start()].count("\n") + 1
        fp[name + "_matrices"] = {"values": vals, "source": f"tests/qcd/fermion_operators.py:{line}"}
    vec["fingerprints"] = fp
    # ---- Wilson pion correlator (eo2 CG, single precision) --------------------------------------------------------------
    m = re.search(r"correlator_ref\s*=\s*\[(.*?)\]", src, re.S)
    if m:
        vec["wilson_pion_correlator"] = {"values": [float(x) for x in re.findall(r"[-0-9.e]+", m.group(1))],
                                         "source": f"tests/qcd/fermion_operators.py:{src[:m.start()].count(chr(10)) + 1}"}
    # ---- domain wall: Moebius correlator and J5q ----------------------------------------------------------------------------
    f = os.path.join(root, "tests/qcd/domain_wall.py")
    a = _assignments(f)
    for name in ["correlator_ref", "J5q_ref"]:
        vec["domain_wall_" + name] = {"values": list(a[name][0]), "source": f"tests/qcd/domain_wall.py:{a[name][1]}"}
    # ---- random numbers --------------------------------------------------------------------------------------------------
    f = os.path.join(root, "tests/random/simple.py")
    src = open(f).read()
    seqs = []
    for m in re.finditer(r"test_sequence_ref\s*=\s*np\.array\(\s*\[(.*?)\]", src, re.S):
        seqs.append({"values": [float(x) for x in re.findall(r"-?[0-9]+\.[0-9]+(?:e-?[0-9]+)?", m.group(1))],
                     "source": f"tests/random/simple.py:{src[:m.start()].count(chr(10)) + 1}"})
    vec["rng_normal_sequences"] = seqs
    m = re.search(r"for grid, prec, ref_plaquette, scale in \[(.*?)\]:", src, re.S)
    rows = re.findall(r"\(\s*(\w+)\s*,\s*([-0-9.e]+)\s*,\s*([-0-9.e]+)\s*,\s*([-0-9.e]+)\s*\)", m.group(1))
    vec["rng_gauge_random_plaquettes"] = {
        "values": [{"grid": g_, "tolerance2": float(t), "plaquette": float(pq), "scale": float(sc)} for g_, t, pq, sc in rows],
        "source": f"tests/random/simple.py:{src[:m.start()].count(chr(10)) + 1}"}
    m = re.search(r'rng\.choice\(\["A", "B", "C"\], 10\) == \[(.*?)\]', src, re.S)
    vec["rng_choice_letters"] = {"values": re.findall(r'"([ABC])"', m.group(1)), "source": f"tests/random/simple.py:{src[:m.start()].count(chr(10)) + 1}"}
    m = re.search(r"np\.array\(\[1, 2, 3\]\), 5\) == np\.array\(\[(.*?)\]\)", src)
    vec["rng_choice_numbers"] = {"values": [int(x) for x in m.group(1).split(",")], "source": f"tests/random/simple.py:{src[:m.start()].count(chr(10)) + 1}"}
    return vec


if __name__ == "__main__":
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")
    with open(out, "w") as fh:
        json.dump(extract(root), fh, indent=1, sort_keys=True)
    print("wrote", out)
