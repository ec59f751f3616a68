"""Minimal MyHDL-compatible simulation layer (clean-room, API surface only).

The reference (tomtor/HDL-deflate) is written against the external `myhdl`
package (deflate.py:15-16, test_deflate.py:7-9), which is neither vendored by
the reference nor installed in this image.  This module supplies the subset of
that API the reference's engine and test bench use, so that

  * the unchanged reference `deflate.py` can be clocked on the CPU (that is how
    the oracle's golden vectors are produced, see oracle/ref_sim.py), and
  * the unchanged reference `test_deflate.py` can drive the GPU-backed
    `deflate` module of this repo (dropin/deflate.py) through the same ports.

Semantics implemented (what the reference relies on):
  - `Signal` with two-phase update: `.next` assignments are committed at the
    end of the current delta cycle; `.next` read-back returns the pending value.
  - `intbv` (range-checked on commit) and `modbv` (wraps on commit), created
    with `intbv(v)[n:]` / `intbv(min=, max=)`.
  - `@always(sig.posedge)`, `@always_comb`, `@instance`, `@block`, `instances()`.
  - `Simulation(*blocks_and_generators).run()` with `yield delay(n)`: values
    written by a generator before its `yield` become visible together with the
    clock edge it toggled (same delta), exactly what test_deflate.py:95-134
    depends on.
  - `concat`, `ConcatSignal`, `enum`, `ResetSignal`, `Error`, `StopSimulation`,
    `now()`.  `block.convert()` is a no-op (HDL generation is out of scope).
"""

import heapq
import inspect
import sys

__all__ = [
    "Signal", "intbv", "modbv", "always", "always_comb", "always_seq", "block",
    "instances", "instance", "enum", "concat", "ConcatSignal", "ResetSignal",
    "Error", "StopSimulation", "Simulation", "Cosimulation", "delay", "now",
    "posedge", "negedge",
]


class Error(Exception):
    pass


class StopSimulation(Exception):
    pass


# --------------------------------------------------------------------------
# bit vectors
# --------------------------------------------------------------------------

def _toint(v):
    if isinstance(v, int):
        return v
    return int(v)


class intbv(object):
    """Integer with an optional [min, max) range / bit width."""

    __slots__ = ("_val", "_min", "_max", "_nrbits")
    _wrap = False

    def __init__(self, val=0, min=None, max=None, _nrbits=0):
        self._val = _toint(val)
        self._min = min
        self._max = max
        if _nrbits:
            self._nrbits = _nrbits
        elif min is not None and max is not None:
            lo, hi = min, max - 1
            if lo >= 0:
                self._nrbits = hi.bit_length()
            else:
                self._nrbits = 1 + builtins_max((-lo - 1).bit_length(),
                                                hi.bit_length() if hi > 0 else 0)
        else:
            self._nrbits = 0

    # -- construction helpers ------------------------------------------------
    def __getitem__(self, key):
        if isinstance(key, slice):
            hi, lo = key.start, key.stop
            lo = 0 if lo is None else _toint(lo)
            if hi is None:
                return type(self)(self._val >> lo)
            hi = _toint(hi)
            n = hi - lo
            if n <= 0:
                raise ValueError("intbv slice: left index must exceed right index")
            v = (self._val >> lo) & ((1 << n) - 1)
            return type(self)(v, min=0, max=1 << n, _nrbits=n)
        i = _toint(key)
        return bool((self._val >> i) & 1)

    def _coerce(self, v):
        """Return the integer a Signal of this type stores for `v`."""
        v = _toint(v)
        if self._wrap:
            if self._nrbits:
                lo = self._min if self._min is not None else 0
                span = 1 << self._nrbits
                if self._min is not None and self._max is not None:
                    span = self._max - self._min
                v = (v - lo) % span + lo
            return v
        if self._max is not None and v >= self._max:
            raise ValueError("intbv value %d >= maximum %d" % (v, self._max))
        if self._min is not None and v < self._min:
            raise ValueError("intbv value %d < minimum %d" % (v, self._min))
        return v

    def __len__(self):
        return self._nrbits

    def __int__(self):
        return self._val

    __index__ = __int__

    def __bool__(self):
        return self._val != 0

    def __repr__(self):
        return "%s(%d)" % (type(self).__name__, self._val)

    # arithmetic on raw intbv objects (rare in the reference): behave as int
    def __add__(self, o): return self._val + _toint(o)
    def __radd__(self, o): return _toint(o) + self._val
    def __sub__(self, o): return self._val - _toint(o)
    def __rsub__(self, o): return _toint(o) - self._val
    def __mul__(self, o): return self._val * _toint(o)
    def __rmul__(self, o): return _toint(o) * self._val
    def __and__(self, o): return self._val & _toint(o)
    def __rand__(self, o): return _toint(o) & self._val
    def __or__(self, o): return self._val | _toint(o)
    def __ror__(self, o): return _toint(o) | self._val
    def __xor__(self, o): return self._val ^ _toint(o)
    def __rxor__(self, o): return _toint(o) ^ self._val
    def __lshift__(self, o): return self._val << _toint(o)
    def __rlshift__(self, o): return _toint(o) << self._val
    def __rshift__(self, o): return self._val >> _toint(o)
    def __rrshift__(self, o): return _toint(o) >> self._val
    def __floordiv__(self, o): return self._val // _toint(o)
    def __mod__(self, o): return self._val % _toint(o)
    def __neg__(self): return -self._val
    def __invert__(self): return ~self._val
    def __eq__(self, o): return self._val == o
    def __ne__(self, o): return self._val != o
    def __lt__(self, o): return self._val < o
    def __le__(self, o): return self._val <= o
    def __gt__(self, o): return self._val > o
    def __ge__(self, o): return self._val >= o
    __hash__ = None


class modbv(intbv):
    __slots__ = ()
    _wrap = True


import builtins as _builtins
builtins_max = _builtins.max


def _bitlen(v):
    return v.bit_length() if v >= 0 else (-v - 1).bit_length() + 1


# --------------------------------------------------------------------------
# enum
# --------------------------------------------------------------------------

class _EnumItem(object):
    __slots__ = ("_name", "_index", "_owner")

    def __init__(self, name, index, owner):
        self._name, self._index, self._owner = name, index, owner

    def __repr__(self):
        return self._name

    def __int__(self):
        return self._index

    __index__ = __int__

    def __hash__(self):
        return id(self)

    def __eq__(self, o):
        if isinstance(o, _SignalOps):
            o = o.val
        return self is o

    def __ne__(self, o):
        return not self.__eq__(o)


class _Enum(object):
    def __init__(self, names):
        self._names = names
        self._items = []
        for i, n in enumerate(names):
            it = _EnumItem(n, i, self)
            self._items.append(it)
            setattr(self, n, it)

    def __len__(self):
        return len(self._items)


def enum(*names, **kw):
    return _Enum(names)


# --------------------------------------------------------------------------
# signals
# --------------------------------------------------------------------------

_pending = []          # signals with an uncommitted .next
_tracking = None       # set() collecting signals read by an always_comb body


class _Edge(object):
    __slots__ = ("sig", "rising")

    def __init__(self, sig, rising):
        self.sig, self.rising = sig, rising


def posedge(sig):
    return sig.posedge


def negedge(sig):
    return sig.negedge


class _SignalOps(object):
    """Integer-like behaviour shared by Signal and ConcatSignal (uses .val)."""

    __slots__ = ()

    def __int__(self): return int(self.val)
    def __index__(self): return int(self.val)
    def __bool__(self): return bool(self.val)
    def __len__(self): return self._nrbits

    def __add__(self, o): return self.val + _v(o)
    def __radd__(self, o): return _v(o) + self.val
    def __sub__(self, o): return self.val - _v(o)
    def __rsub__(self, o): return _v(o) - self.val
    def __mul__(self, o): return self.val * _v(o)
    def __rmul__(self, o): return _v(o) * self.val
    def __floordiv__(self, o): return self.val // _v(o)
    def __rfloordiv__(self, o): return _v(o) // self.val
    def __mod__(self, o): return self.val % _v(o)
    def __rmod__(self, o): return _v(o) % self.val
    def __and__(self, o): return self.val & _v(o)
    def __rand__(self, o): return _v(o) & self.val
    def __or__(self, o): return self.val | _v(o)
    def __ror__(self, o): return _v(o) | self.val
    def __xor__(self, o): return self.val ^ _v(o)
    def __rxor__(self, o): return _v(o) ^ self.val
    def __lshift__(self, o): return self.val << _v(o)
    def __rlshift__(self, o): return _v(o) << self.val
    def __rshift__(self, o): return self.val >> _v(o)
    def __rrshift__(self, o): return _v(o) >> self.val
    def __neg__(self): return -self.val
    def __pos__(self): return +self.val
    def __invert__(self): return ~self.val
    def __abs__(self): return abs(self.val)

    def __eq__(self, o): return self.val == _v(o)
    def __ne__(self, o): return self.val != _v(o)
    def __lt__(self, o): return self.val < _v(o)
    def __le__(self, o): return self.val <= _v(o)
    def __gt__(self, o): return self.val > _v(o)
    def __ge__(self, o): return self.val >= _v(o)

    def __getitem__(self, key):
        v = int(self.val)
        if isinstance(key, slice):
            hi, lo = key.start, key.stop
            lo = 0 if lo is None else _toint(lo)
            if hi is None:
                return v >> lo
            return (v >> lo) & ((1 << (_toint(hi) - lo)) - 1)
        return bool((v >> _toint(key)) & 1)

    def _markUsed(self):
        pass

    def _markRead(self):
        pass


def _v(o):
    if isinstance(o, _SignalOps):
        return o.val
    if isinstance(o, intbv):
        return o._val
    return o


class Signal(_SignalOps):
    __slots__ = ("_cur", "_nxt", "_has_next", "_type", "_nrbits", "_pos", "_neg",
                 "_comb", "_posedge", "_negedge", "_isbool", "__weakref__")

    def __init__(self, val=None, delay=None):
        if isinstance(val, intbv):
            self._type = val
            self._cur = val._val
            self._nrbits = val._nrbits
            self._isbool = False
        elif isinstance(val, bool):
            self._type = None
            self._cur = val
            self._nrbits = 1
            self._isbool = True
        else:
            self._type = None
            self._cur = val
            self._nrbits = 0
            self._isbool = False
        self._nxt = self._cur
        self._has_next = False
        self._pos = []       # processes waiting on posedge
        self._neg = []
        self._comb = []      # always_comb processes reading this signal
        self._posedge = None
        self._negedge = None

    # current value ---------------------------------------------------------
    @property
    def val(self):
        if _tracking is not None:
            _tracking.add(self)
        return self._cur

    # next value ------------------------------------------------------------
    @property
    def next(self):
        return self._nxt

    @next.setter
    def next(self, v):
        t = self._type
        if t is not None:
            v = t._coerce(_v(v))
        else:
            v = _v(v)
            if self._isbool:
                if isinstance(v, intbv):
                    v = v._val
                if v not in (0, 1):
                    raise ValueError("Expected boolean value, got %r" % (v,))
        self._nxt = v
        if not self._has_next:
            self._has_next = True
            _pending.append(self)

    @property
    def posedge(self):
        if self._posedge is None:
            self._posedge = _Edge(self, True)
        return self._posedge

    @property
    def negedge(self):
        if self._negedge is None:
            self._negedge = _Edge(self, False)
        return self._negedge

    @property
    def min(self):
        return self._type._min if self._type is not None else None

    @property
    def max(self):
        return self._type._max if self._type is not None else None

    def __hash__(self):
        return id(self)

    def __repr__(self):
        return "Signal(%r)" % (self._cur,)


class ResetSignal(Signal):
    __slots__ = ("active", "isasync")

    def __init__(self, val, active, isasync=None, **kw):
        Signal.__init__(self, bool(val))
        self.active = active
        self.isasync = kw.get("async", isasync)


class ConcatSignal(_SignalOps):
    """Read-only view: the concatenation (MSB first) of its argument signals."""

    __slots__ = ("_args", "_nrbits")

    def __init__(self, *args):
        self._args = args
        n = 0
        for a in args:
            w = len(a)
            if not w:
                raise ValueError("ConcatSignal arguments need a bit width")
            n += w
        self._nrbits = n

    @property
    def val(self):
        r = 0
        for a in self._args:
            r = (r << len(a)) | (int(a.val) & ((1 << len(a)) - 1))
        return r

    def __hash__(self):
        return id(self)


def concat(base, *args):
    """Concatenate `base` (any width) with sized arguments, MSB first."""
    r = int(_v(base))
    for a in args:
        if isinstance(a, bool):
            w, av = 1, int(a)
        elif isinstance(a, str):
            w, av = len(a), int(a, 2)
        else:
            w = len(a)
            av = int(_v(a))
            if not w:
                raise ValueError("concat argument without a bit width")
        r = (r << w) | (av & ((1 << w) - 1))
    return r


# --------------------------------------------------------------------------
# processes and blocks
# --------------------------------------------------------------------------

class _Process(object):
    __slots__ = ("func", "edges", "kind", "gen", "_seen")

    def __init__(self, func, kind, edges=()):
        self.func = func
        self.kind = kind          # 'seq' | 'comb' | 'instance'
        self.edges = edges
        self.gen = None
        self._seen = -1


def always(*events):
    edges = []
    for e in events:
        if isinstance(e, _Edge):
            edges.append(e)
        elif isinstance(e, Signal):
            edges.append(e.posedge)
            edges.append(e.negedge)
        elif isinstance(e, delay):
            edges.append(e)
        else:
            raise TypeError("always(): unsupported event %r" % (e,))

    def deco(func):
        return _Process(func, "seq", tuple(edges))
    return deco


def always_seq(edge, reset=None):
    def deco(func):
        return _Process(func, "seq", (edge,))
    return deco


def always_comb(func):
    return _Process(func, "comb")


def instance(genfunc):
    return _Process(genfunc, "instance")


class _Block(object):
    def __init__(self, func, args, kwargs):
        self.func = func
        self.name = func.__name__
        subs = func(*args, **kwargs)
        self.subs = _flatten(subs)

    def convert(self, *a, **kw):      # HDL generation: out of scope
        return None

    def config_sim(self, *a, **kw):
        return None

    def verify_convert(self, *a, **kw):
        return None

    def run_sim(self, duration=None, quiet=0):
        Simulation(self).run(duration, quiet)

    def _processes(self, out):
        for s in self.subs:
            if isinstance(s, _Block):
                s._processes(out)
            else:
                out.append(s)
        return out


def _flatten(x):
    out = []
    if x is None:
        return out
    if isinstance(x, (_Process, _Block)):
        return [x]
    if isinstance(x, (list, tuple, set)):
        for y in x:
            out.extend(_flatten(y))
        return out
    if inspect.isgenerator(x):
        p = _Process(None, "instance")
        p.gen = x
        return [p]
    raise TypeError("block returned an unsupported object: %r" % (x,))


def block(func):
    def make(*args, **kwargs):
        return _Block(func, args, kwargs)
    make.__name__ = func.__name__
    make.__doc__ = func.__doc__
    make._is_block_factory = True
    return make


def instances():
    """Collect processes / blocks (and lists of them) from the caller's locals."""
    f = sys._getframe(1)
    out = []
    for v in f.f_locals.values():
        if isinstance(v, (_Process, _Block)):
            out.append(v)
        elif isinstance(v, (list, tuple)) and v and \
                all(isinstance(y, (_Process, _Block)) for y in v):
            out.extend(v)
    return out


# --------------------------------------------------------------------------
# simulator
# --------------------------------------------------------------------------

class delay(object):
    __slots__ = ("t",)

    def __init__(self, t):
        self.t = int(t)


_now = 0
_stamp = 0


def now():
    return _now


class Cosimulation(object):
    def __init__(self, *a, **kw):
        raise Error("Cosimulation is not available in the compat layer")


class Simulation(object):
    def __init__(self, *args):
        procs = []
        for a in _flatten(list(args)):
            if isinstance(a, _Block):
                a._processes(procs)
            else:
                procs.append(a)
        self._seq = [p for p in procs if p.kind == "seq"]
        self._comb = [p for p in procs if p.kind == "comb"]
        self._inst = [p for p in procs if p.kind == "instance"]
        self._heap = []
        self._order = 0
        self._started = False

    # -- wiring ---------------------------------------------------------------
    def _start(self):
        global _now
        _now = 0
        for p in self._seq:
            for e in p.edges:
                if isinstance(e, _Edge):
                    (e.sig._pos if e.rising else e.sig._neg).append(p)
                else:
                    raise Error("always(delay) is not supported by the compat layer")
        for p in self._inst:
            if p.gen is None:
                p.gen = p.func()
            self._push(0, p)
        for p in self._comb:
            self._run_comb(p)
        self._started = True

    def _push(self, t, p):
        self._order += 1
        heapq.heappush(self._heap, (t, self._order, p))

    def _run_comb(self, p):
        global _tracking
        _tracking = s = set()
        try:
            p.func()
        finally:
            _tracking = None
        for sig in s:
            if p not in sig._comb:
                sig._comb.append(p)

    # -- delta cycles ---------------------------------------------------------
    def _settle(self):
        global _pending, _stamp
        while _pending:
            batch = _pending
            _pending = []
            seq_run = []
            comb_run = []
            _stamp += 1
            stamp = _stamp
            for s in batch:
                s._has_next = False
                new, old = s._nxt, s._cur
                if new == old and type(new) is type(old):
                    continue
                s._cur = new
                if s._comb:
                    for p in s._comb:
                        if p._seen != (stamp, 0):
                            p._seen = (stamp, 0)
                            comb_run.append(p)
                if s._pos and not old and new:
                    for p in s._pos:
                        if p._seen != (stamp, 1):
                            p._seen = (stamp, 1)
                            seq_run.append(p)
                if s._neg and old and not new:
                    for p in s._neg:
                        if p._seen != (stamp, 1):
                            p._seen = (stamp, 1)
                            seq_run.append(p)
            for p in seq_run:
                p.func()
            for p in comb_run:
                self._run_comb(p)

    # -- main loop ------------------------------------------------------------
    def run(self, duration=None, quiet=0):
        global _now
        if not self._started:
            self._start()
        limit = None if duration is None else _now + duration
        try:
            self._settle()
            while self._heap:
                t, _, p = self._heap[0]
                if limit is not None and t > limit:
                    _now = limit
                    return 1
                _now = t
                # run every generator scheduled for this instant
                while self._heap and self._heap[0][0] == t:
                    _, _, p = heapq.heappop(self._heap)
                    try:
                        y = next(p.gen)
                    except StopIteration:
                        continue
                    if isinstance(y, delay):
                        self._push(t + y.t, p)
                    elif isinstance(y, (list, tuple)) and y and isinstance(y[0], delay):
                        self._push(t + y[0].t, p)
                    else:
                        raise Error("compat Simulation: generators may only yield delay()")
                self._settle()
        except StopSimulation:
            return 0
        finally:
            for p in self._seq:
                for e in p.edges:
                    lst = e.sig._pos if e.rising else e.sig._neg
                    if p in lst:
                        lst.remove(p)
        return 0

    def quit(self):
        pass


def traceSignals(dut, *a, **kw):
    return dut
