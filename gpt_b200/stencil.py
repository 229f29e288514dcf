"""g.stencil.matrix_vector (lib/gpt/core/stencil/matrix_vector.py:140-175): the distributed front end of the local stencil.  On
one rank -- what this package's executor covers -- both are the same object."""
from gpt_b200 import local_stencil


def matrix_vector(lat_matrix, lat_vector, points, code, code_parallel_block_size=None, matrix_parity=0, vector_parity=0):
    return local_stencil.matrix_vector(lat_matrix, lat_vector, points, code, code_parallel_block_size, local=0,
                                       matrix_parity=matrix_parity, vector_parity=vector_parity)
