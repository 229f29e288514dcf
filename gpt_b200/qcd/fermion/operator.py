"""
Fermion-operator front end: parameter normalisation, grids, opcode table, handle cache.
Mirrors lib/gpt/qcd/fermion/operator/{base,fine_operator,interface}.py of the reference:

  interface   operator/interface.py:28-100  (handle cache `operator_limbo`, (src,dst) call order)
  base        operator/base.py:28-304       (grids, matrix_operator wrappers, adj/updated/converted/modified,
                                             propagator, even_odd_sites_decomposed)
"""
import gpt_b200 as g
import cgpt
from gpt_b200 import capi
from gpt_b200.qcd.fermion.register import register

operator_tag = {}
operator_limbo = {}


class interface:
    def __init__(self):
        self.obj = None

    def _setup(self, name, grid, params):
        assert self.obj is None
        tag_params = {x: params[x] for x in params if x not in ["U", "mass", "mass_plus", "mass_minus"]}
        # the reference's params carry the grid handles (U_grid, F_grid, ...), which makes the tag grid specific
        tag = f"{name}_{grid.precision.cgpt_dtype}_{list(grid.fdimensions)}_{getattr(grid, 'mpi', None)}_{tag_params}"
        if tag in operator_limbo and len(operator_limbo[tag]) > 0:
            self.obj = operator_limbo[tag].pop()
            # set mass needs to precede update for clover-type fermions (interface.py:43-45)
            cgpt.set_mass_fermion_operator(self.obj, params)
            cgpt.update_fermion_operator(self.obj, params)
        else:
            self.obj = cgpt.create_fermion_operator(name, grid.precision.cgpt_dtype, params)
            operator_tag[self.obj] = tag

    def setup(self, name, grid, params):
        self.setup_arguments = (name, grid, params)
        self._setup(*self.setup_arguments)

    def __del__(self):
        self.suspend()

    def suspend(self):
        if self.obj is not None:
            tag = operator_tag[self.obj]
            operator_limbo.setdefault(tag, []).append(self.obj)
            self.obj = None

    def update(self, params):
        if self.obj is None:
            self._setup(*self.setup_arguments)
        cgpt.update_fermion_operator(self.obj, params)

    def apply_unary_operator_host(self, opcode, o, i):
        """o, i: C-contiguous numpy arrays (or anything exposing __array_interface__ / data_ptr) holding full fields in
        GPT order; equivalent to  lattice[:] = i ; apply ; o[:] = lattice[:]  with the copies overlapped"""
        assert self.obj is not None
        return capi.apply_fermion_operator_host(self.obj, opcode, _address(i), _address(o), _nbytes(i))

    def apply_unary_operator(self, opcode, o, i):
        assert self.obj is not None
        # cgpt adopts Grid's (in, out) order (interface.py:88-91)
        return cgpt.apply_fermion_operator(self.obj, opcode, i.v_obj, o.v_obj)


def _address(a):
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return a.__array_interface__["data"][0]


def _nbytes(a):
    if hasattr(a, "data_ptr"):
        return a.numel() * a.element_size()
    return a.nbytes


class method_registry:
    pass


class fine_operator(g.matrix_operator):
    def __init__(self, name, U, params, otype=None, daggered=False):
        self.name = name
        self.U = U
        self.otype = otype if otype is not None else g.ot_vector_spin_color(4, 3)
        self.params_constructor = params
        self.daggered = daggered

        self.U_grid = U[0].grid
        self.U_grid_eo = g.grid(self.U_grid.fdimensions, self.U_grid.precision, g.redblack)
        if "Ls" not in params or params["Ls"] is None:
            self.F_grid = self.U_grid
            self.F_grid_eo = self.U_grid_eo
        else:
            self.F_grid = self.U_grid.inserted_dimension(0, params["Ls"])
            self.F_grid_eo = g.grid(self.F_grid.fdimensions, self.U_grid.precision, g.redblack)

        # the grid handles travel with the parameters like in the reference (operator/base.py:65-76): they make the handle-cache
        # tag of interface.py grid specific
        self.params = {"U_grid": self.U_grid.obj, "U_grid_rb": self.U_grid_eo.obj, "F_grid": self.F_grid.obj,
                       "F_grid_rb": self.F_grid_eo.obj, "U": [u.obj for u in U]}
        for k in params:
            assert k not in ["U_grid", "U_grid_rb", "F_grid", "F_grid_rb", "U"]
            self.params[k] = params[k]

        self.interface = interface()
        self.interface.setup(name, self.U_grid, self.params)

        registry = method_registry()
        register(registry, self.interface)

        otype = self.otype
        self.vector_space_F = g.vector_space(self.F_grid, otype)
        self.vector_space_U = g.vector_space(self.U_grid, otype)
        self.vector_space_F_eo = g.vector_space(self.F_grid_eo, otype)

        super().__init__(
            mat=registry.M if not daggered else registry.Mdag,
            adj_mat=registry.Mdag if not daggered else registry.M,
            vector_space=self.vector_space_F,
        )

        def OP(x):
            return x if not daggered else x.adj()

        mo = g.matrix_operator
        self.Meooe = OP(mo(mat=registry.Meooe, adj_mat=registry.MeooeDag, vector_space=self.vector_space_F_eo))
        self.Mooee = OP(
            mo(mat=registry.Mooee, adj_mat=registry.MooeeDag, inv_mat=registry.MooeeInv,
               adj_inv_mat=registry.MooeeInvDag, vector_space=self.vector_space_F_eo)
        )
        self.DhopEO = OP(mo(mat=registry.DhopEO, adj_mat=registry.DhopEODag, vector_space=self.vector_space_F_eo))
        self.Mdiag = OP(mo(registry.Mdiag, vector_space=self.vector_space_F))
        self.Dminus = OP(mo(mat=registry.Dminus, adj_mat=registry.DminusDag, vector_space=self.vector_space_F))
        self.ImportPhysicalFermionSource = OP(
            mo(registry.ImportPhysicalFermionSource, vector_space=(self.vector_space_F, self.vector_space_U))
        )
        self.ImportUnphysicalFermion = OP(
            mo(registry.ImportUnphysicalFermion, vector_space=(self.vector_space_F, self.vector_space_U))
        )
        self.ExportPhysicalFermionSolution = OP(
            mo(registry.ExportPhysicalFermionSolution, vector_space=(self.vector_space_U, self.vector_space_F))
        )
        self.ExportPhysicalFermionSource = OP(
            mo(registry.ExportPhysicalFermionSource, vector_space=(self.vector_space_U, self.vector_space_F))
        )
        self.Dhop = OP(mo(mat=registry.Dhop, adj_mat=registry.DhopDag, vector_space=self.vector_space_F))
        # host-buffer variant of Dhop: op.Dhop_host(dst_array, src_array)
        code = 3001 if not daggered else 4001
        iface = self.interface  # not `self`: the operator must not hold a reference to itself (fermion_operators.py:866-872)
        self.Dhop_host = lambda dst, src: iface.apply_unary_operator_host(code, dst, src)

    # -- variations (base.py:222-262)
    def modified(self, **params):
        return type(self)(self.name, self.U, {**self.params_constructor, **params}, self.otype, self.daggered)

    def converted(self, dst_precision):
        return self.updated(g.convert(self.U, dst_precision))

    def updated(self, U):
        return type(self)(self.name, U, self.params_constructor, self.otype, self.daggered)

    def adj(self):
        return type(self)(self.name, self.U, self.params_constructor, self.otype, not self.daggered)

    def update(self, U):
        self.U = U
        self.params["U"] = [u.obj for u in U]
        self.interface.update(self.params)

    def suspend(self):
        self.interface.suspend()

    def arguments(self):
        return self.U

    # -- propagator (base.py:270-286)
    def propagator(self, solver):
        exp = self.ExportPhysicalFermionSolution
        imp = self.ImportPhysicalFermionSource
        inv_matrix = solver(self)

        def prop(dst_sc, src_sc):
            g.eval(dst_sc, exp * inv_matrix * imp * g.expr(src_sc))

        op = g.matrix_operator(prop, vector_space=(exp.vector_space[0], imp.vector_space[1]))
        if self.daggered:
            op = op.adj()
        return g.propagator_operator(op)

    # -- even/odd decomposition (base.py:288-309)
    def even_odd_sites_decomposed(self, parity):
        me_op = self

        class even_odd_sites:
            def __init__(me):
                me.op = me_op
                me.parity = parity
                me.DD = me_op.Mooee.clone()
                me.CC = me_op.Mooee.clone()
                me.CD = me_op.Meooe.clone()
                me.DC = me_op.Meooe.clone()
                me.DD.vector_space[1].cb = parity
                me.DD.vector_space[0].cb = parity
                me.CC.vector_space[1].cb = parity.inv()
                me.CC.vector_space[0].cb = parity.inv()
                me.CD.vector_space[1].cb = parity
                me.CD.vector_space[0].cb = parity.inv()
                me.DC.vector_space[1].cb = parity.inv()
                me.DC.vector_space[0].cb = parity

            # D_domain / C_domain project & promote (lib/gpt/core/domain/even_odd_sites.py)
            def project(me, cb, full_field):
                half = g.lattice(me_op.F_grid_eo, me_op.otype)
                g.pick_checkerboard(cb, half, full_field)
                return half

            def promote(me, full_field, half):
                g.set_checkerboard(full_field, half)

        return even_odd_sites()


# g.qcd.fermion.operator.base.base (lib/gpt/qcd/fermion/operator/base.py): the class user code tests operators against
import types as _types  # noqa: E402

base = _types.SimpleNamespace(base=fine_operator)


class differentiable_fine_operator(fine_operator):
    """operators with projected-gradient (MDeriv...) entry points (lib/gpt/qcd/fermion/operator/differentiable_fine_operator.py): the
    force terms of HMC are outside the hot path, no operator of this package is an instance"""


class gauge_independent_g5_hermitian:
    """marker class of lib/gpt/qcd/fermion/operator/fine_operator.py (G5 hermiticity of Wilson-type operators)"""
