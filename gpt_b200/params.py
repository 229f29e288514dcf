"""@params_convention of GPT (lib/gpt/params.py:20-82): the decorated function takes its parameters as a dict in the LAST
argument; callers may pass dicts and/or keyword arguments, defaults are filled in, positional arguments with defaults may be
omitted, unknown keys are rejected."""
import inspect


def params_convention(default=None, **kwdefault):
    defaults = {**(default or {}), **kwdefault}

    def decorate(f):
        fparams = list(inspect.signature(f).parameters.values())
        assert len(fparams) > 0
        nargs = len(fparams) - 1  # the last one is the params dict
        nargs_min = nargs
        for i in reversed(range(nargs)):
            if fparams[i].default is inspect.Parameter.empty:
                nargs_min = i + 1
                break
        else:
            nargs_min = 0

        def wrap(*args, **kwargs):
            assert len(args) >= nargs_min
            for i in range(len(args), nargs):
                args = args + (fparams[i].default,)
            positional = args[:nargs]
            params = {**{k: v for d in args[nargs:] for k, v in d.items()}, **kwargs}
            for k in params:
                if k not in defaults:
                    raise KeyError(f"Parameter {k} is not known (allowed: {sorted(defaults)})")
            return f(*positional, {**defaults, **params})

        wrap.__name__ = getattr(f, "__name__", "wrap")
        wrap.__doc__ = f.__doc__
        return wrap

    return decorate
