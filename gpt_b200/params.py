"""@params_convention: GPT's calling convention for parameter dictionaries (semantics of lib/gpt/params.py:20-82).

A decorated function receives all of its named parameters as ONE dict in its last argument.  Callers may hand over any mix
of dicts (after the positional arguments) and keyword arguments; missing keys take the declared defaults, positional
arguments that have defaults of their own may be left out, and a key that was not declared is an error.
"""
import inspect


def params_convention(default=None, **kwdefault):
    declared = dict(default or {})
    declared.update(kwdefault)

    def decorate(func):
        signature = list(inspect.signature(func).parameters.values())
        if not signature:
            raise TypeError("a params_convention function needs a last argument that receives the parameter dict")
        positional = signature[:-1]
        # the leading positional arguments without a default are mandatory
        n_required = 0
        for i, prm in enumerate(positional):
            if prm.default is inspect.Parameter.empty:
                n_required = i + 1

        def wrapper(*args, **kwargs):
            if len(args) < n_required:
                raise TypeError(f"{getattr(func, '__name__', 'function')} needs at least {n_required} positional arguments")
            head = list(args[:len(positional)])
            head += [prm.default for prm in positional[len(head):]]
            merged = {}
            for extra in args[len(positional):]:
                merged.update(extra)
            merged.update(kwargs)
            unknown = [k for k in merged if k not in declared]
            if unknown:
                raise KeyError(f"Parameter {unknown[0]} is not known (allowed: {sorted(declared)})")
            return func(*head, {**declared, **merged})

        wrapper.__name__ = getattr(func, "__name__", "wrapper")
        wrapper.__doc__ = func.__doc__
        return wrapper

    return decorate
