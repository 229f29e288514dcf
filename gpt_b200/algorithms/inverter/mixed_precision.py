"""inv.mixed_precision (lib/gpt/algorithms/inverter/mixed_precision.py:23-32)"""
import gpt_b200 as g
from gpt_b200.algorithms.base import base


class mixed_precision(base):
    def __init__(self, inverter, inner_precision, outer_precision):
        super().__init__()
        self.inverter = inverter
        self.inner_precision = inner_precision
        self.outer_precision = outer_precision

    def __call__(self, mat):
        matrix = mat.converted(self.inner_precision)
        inner = self.inverter(matrix)
        outer_vs = mat.vector_space

        # matrix_operator.converted(outer_precision): convert in, apply, convert out
        # (lib/gpt/core/operator/matrix_operator.py:139-178)
        def inv(dst, src):
            s = g.convert(src, self.inner_precision)
            d = g.convert(dst, self.inner_precision)
            inner(d, s)
            g.convert(dst, d)

        return g.matrix_operator(mat=inv, inv_mat=mat, vector_space=outer_vs, accept_guess=(True, False))
