"""name -> opcode table of the fermion-operator boundary (lib/gpt/qcd/fermion/register.py:2-34 is generated
from lib/cgpt/lib/operators/types.h:26-63; the unary codes are the same here)."""

OPCODES = {
    "M": 2001, "Mdag": 2002, "Meooe": 2003, "MeooeDag": 2004, "Mooee": 2005, "MooeeDag": 2006,
    "MooeeInv": 2007, "MooeeInvDag": 2008, "Mdiag": 2009, "Dminus": 2010, "DminusDag": 2011,
    "ImportPhysicalFermionSource": 2012, "ImportUnphysicalFermion": 2013,
    "ExportPhysicalFermionSolution": 2014, "ExportPhysicalFermionSource": 2015,
    "Dhop": 3001, "DhopDag": 4001, "DhopEO": 3002, "DhopEODag": 4002,
}


def register(reg, op):
    for name, code in OPCODES.items():
        setattr(reg, name, (lambda c: lambda dst, src: op.apply_unary_operator(c, dst, src))(code))
