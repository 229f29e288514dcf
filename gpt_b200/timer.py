"""
g.timer -- sectioned stopwatch of GPT scripts (behaviour of lib/gpt/core/time.py:252-330):

    t = g.timer("Full Timings")
    for n in range(N):
        t("Dhop", flops, nbytes)     # closes the running section, opens "Dhop" (flop / byte counts are optional)
        ...
        t("Mooee")
        ...
        t()                          # closes the running section
    g.message(t)                     # table: total, share, calls, min / mean / max, GFlop/s, GB/s per section

Sections are timed on the HOST clock after a device barrier, like the reference, which synchronises inside every cgpt call:
the library here enqueues asynchronously, so a section boundary drains the stream first.  `t += other` merges two timers.
"""
from gpt_b200 import capi


class _section:
    __slots__ = ("total", "calls", "shortest", "longest", "flop", "byte")

    def __init__(self):
        self.total, self.calls, self.shortest, self.longest, self.flop, self.byte = 0.0, 0, None, None, None, None

    def add(self, dt, flop, byte):
        self.total += dt
        self.calls += 1
        self.shortest = dt if self.shortest is None else min(self.shortest, dt)
        self.longest = dt if self.longest is None else max(self.longest, dt)
        if flop is not None:
            self.flop = (self.flop or 0.0) + flop
        if byte is not None:
            self.byte = (self.byte or 0.0) + byte

    def merge(self, o):
        if o.calls == 0:
            return
        self.total += o.total
        self.calls += o.calls
        self.shortest = o.shortest if self.shortest is None else min(self.shortest, o.shortest)
        self.longest = o.longest if self.longest is None else max(self.longest, o.longest)
        for name in ("flop", "byte"):
            if getattr(o, name) is not None:
                setattr(self, name, (getattr(self, name) or 0.0) + getattr(o, name))


class timer:
    def __init__(self, name="", enabled=True):
        self.name, self.enabled = name, enabled
        self.reset()

    def reset(self):
        self.sections = {}
        self._open = None  # (tag, t0, flop, byte)

    @staticmethod
    def _now():
        from gpt_b200.core import time

        try:
            capi.accelerator_barrier()
        except RuntimeError:
            pass  # no device yet (a timer around pure host code)
        return time()

    def __call__(self, tag=None, flop=None, byte=None):
        if not self.enabled:
            return
        now = self._now()
        if self._open is not None:
            name, t0, fl, by = self._open
            self.sections.setdefault(name, _section()).add(now - t0, fl, by)
        self._open = None if tag is None else (tag, now, flop, byte)

    def __iadd__(self, other):
        for tag, sec in other.sections.items():
            self.sections.setdefault(tag, _section()).merge(sec)
        return self

    def create_point(self):
        return {tag: sec.total for tag, sec in self.sections.items()}

    def __str__(self):
        total = sum(s.total for s in self.sections.values())
        width = max([len(t) for t in self.sections] + [7])
        lines = [f"{self.name}: {total:.3e} s in {len(self.sections)} sections",
                 f"  {'section':<{width}}  {'total/s':>10} {'share':>7} {'calls':>7} {'min/s':>10} {'mean/s':>10} {'max/s':>10} {'GFlop/s':>10} {'GB/s':>10}"]
        for tag, s in sorted(self.sections.items(), key=lambda kv: -kv[1].total):
            gf = f"{s.flop / s.total / 1e9:10.2f}" if s.flop is not None and s.total > 0 else " " * 10
            gb = f"{s.byte / s.total / 1e9:10.2f}" if s.byte is not None and s.total > 0 else " " * 10
            share = 100.0 * s.total / total if total > 0 else 0.0
            lines.append(f"  {tag:<{width}}  {s.total:10.3e} {share:6.2f}% {s.calls:7d} {s.shortest:10.3e} {s.total / s.calls:10.3e} {s.longest:10.3e} {gf} {gb}")
        return "\n".join(lines)
