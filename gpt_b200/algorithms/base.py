"""logging / history plumbing of GPT's algorithms (lib/gpt/algorithms/base.py:23-119)"""
import gpt_b200 as g


class base:
    def __init__(self):
        self.verbose = False
        self.verbose_performance = False
        self.name = self.__class__.__name__

    def log(self, *a):
        if self.verbose:
            g.message(f"{self.name}:", *a)

    def timed_function(self, function):
        def timed(*a):
            t = _timer()
            return function(*a, t)

        return timed


class _timer:
    def __call__(self, *a):
        pass


class base_iterative(base):
    def __init__(self):
        super().__init__()
        self.history = None
        self.verbose_convergence = False
        self.converged = None

    def timed_function(self, function):
        def timed(*a):
            self.history = []
            return function(*a, _timer())

        return timed

    def log_convergence(self, iteration, value, target=None):
        if isinstance(iteration, tuple):
            iteration = iteration[-1]
        self.history.append(value)
        if self.verbose_convergence:
            g.message(f"{self.name}: iteration {iteration}: {value:e} / {target:e}")
