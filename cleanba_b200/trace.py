"""Host-side timeline of the Sebulba pipeline in the Chrome / Perfetto trace-event format (open in ui.perfetto.dev).

The reference's profiling scripts wrap the run in `jax.profiler.trace` and inspect the actor / learner interleaving in
Perfetto (cleanba/cleanba_ppo.py has the same timers as scalars: stats/rollout_time, stats/params_queue_get_time,
stats/rollout_queue_put_time, stats/rollout_queue_get_time, stats/training_time).  `--trace-path run.json` records the same
phases as spans, one track per actor thread plus the learner thread, so the one-version policy lag and who waits on which
queue are visible at a glance.  Spans are host wall-clock (the device work they enqueue is asynchronous unless the phase
itself synchronises, e.g. the per-step action read-back); recording costs two clock reads per span."""
import json
import threading
import time


class _Span:
    __slots__ = ("tr", "name", "tid", "args", "t0")

    def __init__(self, tr, name, tid, args):
        self.tr, self.name, self.tid, self.args = tr, name, tid, args

    def __enter__(self):
        self.t0 = time.perf_counter()
        return self

    def __exit__(self, *exc):
        self.tr._add(self.name, self.tid, self.t0, time.perf_counter(), self.args)
        return False


class Tracer:
    """Thread-safe collector of complete ("X") events; `save()` writes {"traceEvents": [...]}."""

    def __init__(self, process_name: str = "cleanba_b200"):
        self._lock = threading.Lock()
        self._events = []
        self._names = {}
        self._t_origin = time.perf_counter()
        self.process_name = process_name

    def thread_name(self, tid: int, name: str):
        with self._lock:
            self._names[tid] = name

    def span(self, name: str, tid: int, **args):
        return _Span(self, name, tid, args)

    def _add(self, name, tid, t0, t1, args):
        ev = {"name": name, "ph": "X", "pid": 0, "tid": tid, "ts": (t0 - self._t_origin) * 1e6, "dur": (t1 - t0) * 1e6}
        if args:
            ev["args"] = args
        with self._lock:
            self._events.append(ev)

    def events(self):
        with self._lock:
            return list(self._events)

    def save(self, path: str):
        meta = [{"name": "process_name", "ph": "M", "pid": 0, "args": {"name": self.process_name}}]
        with self._lock:
            meta += [{"name": "thread_name", "ph": "M", "pid": 0, "tid": t, "args": {"name": n}} for t, n in sorted(self._names.items())]
            evs = sorted(self._events, key=lambda e: e["ts"])
        with open(path, "w") as f:
            json.dump({"traceEvents": meta + evs, "displayTimeUnit": "ms"}, f)
        return path


class NullTracer:
    """Same interface, records nothing (the default)."""

    class _N:
        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

    _n = _N()

    def thread_name(self, tid, name):
        pass

    def span(self, name, tid, **args):
        return self._n

    def events(self):
        return []

    def save(self, path):
        return None
