"""g.qcd.fermion.reference: the reference ships slow pure-Python restatements of its operators built from covariant shifts
(lib/gpt/qcd/fermion/reference/wilson_clover.py); they exist to cross-check the compiled operators and are not part of the hot
path.  This package's cross-check is oracle/qcd.py (which follows that very file), so the name is kept and calling it says so."""


def wilson_clover(U, params=None, **kwargs):
    raise NotImplementedError("g.qcd.fermion.reference.wilson_clover is the reference's Python cross-check implementation; "
                              "gpt_b200 checks its operators against oracle/qcd.py instead")
