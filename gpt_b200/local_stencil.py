"""g.local_stencil.matrix_vector and g.stencil.matrix_vector (lib/gpt/core/local_stencil/matrix_vector.py:22-74,
lib/gpt/core/stencil/matrix_vector.py:140-175): a list of shifts and a code list, executed by the generic stencil kernel of
libcgpt_b200 (gpt_b200/csrc/stencil.cu) through cgpt.stencil_matrix_vector_*.

Code lines are 6-tuples (target, source, source_point, accumulate, weight, factors) or the dictionaries they stand for;
factors = [(matrix field index, point, adjoint)].  Fields on the full 4d lattice of one rank; points may be any shifts (the
kernel wraps around the lattice itself, so the reference's padded variant for non-cartesian points is the same object here)."""
import cgpt


def parse(c):
    if isinstance(c, tuple):
        assert len(c) == 6
        return {"target": c[0], "source": c[1], "source_point": c[2], "accumulate": c[3], "weight": c[4], "factor": c[5]}
    return c


class matrix_vector:
    def __init__(self, lat_matrix, lat_vector, points, code, code_parallel_block_size=None, local=1, matrix_parity=0, vector_parity=0):
        self.points = points
        self.code = [parse(c) for c in code]
        self.code_parallel_block_size = code_parallel_block_size
        if code_parallel_block_size is None:
            code_parallel_block_size = len(code)
        self.obj = cgpt.stencil_matrix_vector_create(
            lat_matrix.v_obj[0], lat_vector.v_obj[0], lat_matrix.grid.obj, points, self.code, code_parallel_block_size, local,
            matrix_parity, vector_parity)
        self.fast_osites = 0

    def __call__(self, matrix_fields, vector_fields):
        cgpt.stencil_matrix_vector_execute(self.obj, matrix_fields, vector_fields, self.fast_osites)

    def __del__(self):
        if getattr(self, "obj", None) is not None:
            cgpt.stencil_matrix_vector_delete(self.obj)
            self.obj = None

    def data_access_hints(self, *hints):
        pass
