"""
g.default -- command line flags and verbosity switches of a GPT script (API of lib/gpt/default.py:22-92):

    g.default.get_int("--N", 1000), get_float, get_single / get, get_all, get_ivec("--grid", [16,16,16,32], 4), has("--full"),
    is_verbose("cg"), set_verbose("random", False), push_verbose / pop_verbose

Flags are read from sys.argv as `--tag value`; a flag may be repeated (get_all) and get_ivec picks the occurrence whose
dot-separated value has the requested number of components, so `--mpi 1.1.1.2 --mpi 1.1.2` can address 4d and 3d grids.
`--mpi X.Y.Z.T` selects the processor grid of a multi-GPU run (one process per GPU, see gpt_b200/parallel.py).
"""
import sys


def _values_of(tag):
    argv = sys.argv
    return [argv[i + 1] for i in range(len(argv) - 1) if argv[i] == tag]


def get_all(tag, default):
    found = _values_of(tag)
    return found if found else [default]


def has(tag):
    return tag in sys.argv


def get_single(tag, default):
    found = get_all(tag, default)
    if len(found) != 1:
        raise AssertionError(f"flag {tag} given {len(found)} times, expected at most once")
    return found[0]


get = get_single


def _typed(tag, default, convert):
    value = get_single(tag, None)
    return default if value is None else convert(value)


def get_float(tag, default=float("nan")):
    return _typed(tag, default, float)


def get_int(tag, default):
    return _typed(tag, default, int)


def get_ivec(tag, default, ndim):
    for text in _values_of(tag):
        vec = [int(part) for part in text.split(".")]
        if len(vec) == ndim:
            return vec
    return default


# ---- verbosity ---------------------------------------------------------------------------------------------------------
# the solvers of this package log under their class name ("cg", "defect_correcting"); "random" announces rng set-up
_default_on = ["cg", "defect_correcting", "random", "io"]
verbose = set()
_stack = []


def is_verbose(what):
    return what in verbose


def set_verbose(what, status=True):
    (verbose.add if status else verbose.discard)(what)


def push_verbose(what, status):
    _stack.append((what, is_verbose(what)))
    set_verbose(what, status)


def pop_verbose():
    set_verbose(*_stack.pop())


def parse_verbose():
    verbose.clear()
    for group in get_all("--verbose", ",".join(_default_on)):
        verbose.update(x for x in group.split(",") if x)
    for flag, status in (("--verbose_add", True), ("--verbose_remove", False)):
        for group in _values_of(flag):
            for x in group.split(","):
                set_verbose(x, status)


parse_verbose()
