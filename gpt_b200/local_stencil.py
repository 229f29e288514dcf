"""g.local_stencil.matrix_vector and (through gpt_b200/stencil.py) g.stencil.matrix_vector: same call signature as the
reference's lib/gpt/core/local_stencil/matrix_vector.py:22-74 and lib/gpt/core/stencil/matrix_vector.py:140-175, executed by the
generic stencil kernel of libcgpt_b200 (gpt_b200/csrc/stencil.cu) through cgpt.stencil_matrix_vector_*.

A stencil = a list of shifts ("points") + a program.  A program line is the 6-tuple (target, source, source_point, accumulate,
weight, factors) or the dictionary with these keys; factors = [(matrix field index, point, adjoint), ...], applied right to left:

    vector[target](x) = weight * M_1 ... M_n vector[source](x + points[source_point])  [+ vector[accumulate](x)]

Fields live on the full 4d lattice of one rank.  The kernel wraps shifts around the lattice itself, so points may be any shifts
(the reference switches to a padded variant for non-cartesian ones; here it is the same object)."""
import cgpt

_LINE_KEYS = ("target", "source", "source_point", "accumulate", "weight", "factor")


def _program_line(entry):
    line = dict(zip(_LINE_KEYS, entry)) if isinstance(entry, (tuple, list)) else dict(entry)
    missing = [k for k in _LINE_KEYS if k not in line]
    if missing or len(line) != len(_LINE_KEYS):
        raise ValueError(f"stencil code line needs exactly the keys {_LINE_KEYS}, got {sorted(line)}")
    line["factor"] = [tuple(int(v) for v in f) for f in line["factor"]]
    return line


class matrix_vector:
    def __init__(self, lat_matrix, lat_vector, points, code, code_parallel_block_size=None, local=1, matrix_parity=0, vector_parity=0):
        program = [_program_line(entry) for entry in code]
        self.points, self.code = points, program
        self.code_parallel_block_size = code_parallel_block_size
        self.fast_osites = 0  # loop-order hint of the reference's CPU / SIMT back end; the site index is always the fast one here
        self.obj = None
        block = len(program) if code_parallel_block_size is None else int(code_parallel_block_size)
        self.obj = cgpt.stencil_matrix_vector_create(lat_matrix.v_obj[0], lat_vector.v_obj[0], lat_matrix.grid.obj, points, program, block,
                                                     local, matrix_parity, vector_parity)

    def __call__(self, matrix_fields, vector_fields):
        cgpt.stencil_matrix_vector_execute(self.obj, matrix_fields, vector_fields, self.fast_osites)

    def data_access_hints(self, *hints):
        """accepted for compatibility (the padded variant of the reference needs them); nothing to do on one rank"""

    def __del__(self):
        if self.obj is not None:
            cgpt.stencil_matrix_vector_delete(self.obj)
            self.obj = None
