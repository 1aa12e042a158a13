"""`Timer`-compatible statistics with DEVICE time.  Same interface and metric keys as
mineral/common/timer.py:8-85 (`scope / wrap / start / end / stats / reset`; columns count, sum, frac, avg, min, max, total,
totalrate, lastrate), but the reference brackets asynchronous CUDA work with `time.perf_counter()` and never synchronises
(SURVEY.md appendix B item 11), so its "agent.train_epoch" time is the launch time, not the run time.  Here every scope
records a CUDA event pair on the current stream; nothing blocks while the agent runs, and `stats()` synchronises once and reads
the elapsed device time of every finished scope.  Without a CUDA device the scopes fall back to the host clock (the reference's
behaviour) - timing is observability, not part of the data path."""

import collections
import contextlib
import functools
import time

import numpy as np
import torch


class Timer:
    COLUMNS = ('count', 'sum', 'frac', 'avg', 'min', 'max', 'total', 'totalrate', 'lastrate')

    def __init__(self, columns=COLUMNS, device_time=True):
        assert all(x in self.COLUMNS for x in columns), columns
        self.columns = columns
        self.device_time = bool(device_time) and torch.cuda.is_available()
        self._active_scopes = {}
        self._pending = collections.defaultdict(list)    # name -> [(start_event, stop_event)] or [seconds]
        self._totals = collections.defaultdict(float)
        self._last = time.perf_counter()
        self._step = 0

    def reset(self):
        for v in self._active_scopes.values():
            v.__exit__(None, None, None)
        self._active_scopes.clear()
        for timings in self._pending.values():
            timings.clear()
        self._last = time.perf_counter()

    @contextlib.contextmanager
    def scope(self, name):
        if self.device_time:
            start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start.record()
            try:
                yield
            finally:
                stop.record()
                self._pending[name].append((start, stop))
        else:
            t0 = time.perf_counter()
            try:
                yield
            finally:
                self._pending[name].append(time.perf_counter() - t0)

    def wrap(self, name, obj, methods):
        """Times `obj.method` under the scope '<name>.<method>' (timer.py:35-38)."""
        for method in methods:
            fn = getattr(obj, method)
            scope_name = f'{name}.{method}'

            @functools.wraps(fn)
            def timed(*args, __fn=fn, __scope=scope_name, **kwargs):
                with self.scope(__scope):
                    return __fn(*args, **kwargs)

            setattr(obj, method, timed)

    def start(self, name):
        if name in self._active_scopes:
            raise ValueError(f"Scope '{name}' is already active")
        self._active_scopes[name] = self.scope(name)
        self._active_scopes[name].__enter__()

    def end(self, name):
        if name not in self._active_scopes:
            raise ValueError(f"Scope '{name}' is not active.")
        self._active_scopes[name].__exit__(None, None, None)
        del self._active_scopes[name]

    def _durations(self):
        if self.device_time:
            torch.cuda.synchronize()
            return {k: [a.elapsed_time(b) * 1e-3 for a, b in v] for k, v in self._pending.items()}
        return {k: list(v) for k, v in self._pending.items()}

    def stats(self, step=None, reset=True, total_names=None):
        """timer.py:52-85, key for key."""
        durations = self._durations()
        metrics = {'duration': time.perf_counter() - self._last}
        for name, durs in durations.items():
            available = {'count': len(durs), 'sum': float(np.sum(durs))}
            available['frac'] = available['sum'] / metrics['duration']
            if len(durs):
                available['avg'], available['min'], available['max'] = float(np.mean(durs)), float(np.min(durs)), float(np.max(durs))
            self._totals[name] += available['sum']
            available['total'] = self._totals[name]
            if step is not None:
                available['totalrate'] = step / available['total'] if available['total'] else float('inf')
                available['lastrate'] = (step - self._step) / available['sum'] if available['sum'] else float('inf')
            for key, value in available.items():
                if key in self.columns:
                    metrics[f'{name}/{key}'] = value
        if total_names is not None:
            metrics['total'] = sum(self._totals[k] for k in total_names)
            if step is not None:
                metrics['totalrate'] = step / metrics['total'] if metrics['total'] else float('inf')
                metrics['lastrate'] = (step - self._step) / metrics['duration']
        if reset:
            if step is not None:
                self._step = step
            self.reset()
        else:
            self._totals.clear()
        return metrics
