"""@params_convention of GPT (lib/gpt/params.py:22-87): merge a dict argument with keyword arguments on top of
defaults and reject keys that are not declared."""


def params_convention(**defaults):
    def decorate(f):
        def wrap(*args, **kwargs):
            args = list(args)
            given = {}
            if args and isinstance(args[-1], dict):
                given = dict(args.pop())
            given.update(kwargs)
            for k in given:
                if k not in defaults:
                    raise Exception(f"Unknown parameter {k} (allowed: {sorted(defaults)})")
            return f(*args, {**defaults, **given})

        wrap.__name__ = getattr(f, "__name__", "wrap")
        return wrap

    return decorate
