"""
Host-side compatibility layer for ``gymwipe/simtools.py``: ``SimMan`` and ``Notifier``.

The batched simulator replaces the reference's SimPy loop by an event-ordered CUDA kernel; user code
that builds network stacks out of ``Module`` / ``Gate`` / ``Port`` objects (``gymwipe_b200.networking.
construction``) and drives them with generator processes still needs the reference's small process /
event vocabulary on the host -- to be wired, exercised and then TRACED into a scenario table
(``gymwipe_b200.scenario.compile_stack``).  This module provides that vocabulary with the reference's
names and semantics:

* ``SimMan`` (``simtools.py:18-130``): ``init``, ``now``, ``process``, ``event``, ``timeout``,
  ``timeoutUntil``, ``nextTimeSlot``, ``runSimulation`` -- on a small discrete-event engine of its own
  (a heap ordered by ``(time, priority, sequence number)``, the order SimPy 3.0.11 defines: process
  starts are URGENT, everything else NORMAL, ties in insertion order).
* ``Notifier`` (``simtools.py:232-432``): prioritised callbacks, process subscription with the
  ``blocking`` / ``queued`` policies, and the ``event`` property processes wait on.

It is a host utility, not a fallback for the CUDA path: nothing here simulates wireless transmissions.
"""
import heapq
import itertools
from collections import deque

URGENT, NORMAL = 0, 1
_PENDING = object()


class Event:
    """A one-shot event: callbacks run when it is processed; ``value`` is handed to waiting processes."""

    def __init__(self, env):
        self.env = env
        self.callbacks = []
        self._value = _PENDING
        self._ok = True

    @property
    def triggered(self):
        return self._value is not _PENDING

    @property
    def processed(self):
        return self.callbacks is None

    @property
    def ok(self):
        return self._ok

    @property
    def value(self):
        if self._value is _PENDING:
            raise AttributeError("Value of %r is not yet available" % (self,))
        return self._value

    def succeed(self, value=None):
        if self._value is not _PENDING:
            raise RuntimeError("%r has already been triggered" % (self,))
        self._value = value
        self.env._schedule(self, NORMAL, 0.0)
        return self

    def fail(self, exception):
        if self._value is not _PENDING:
            raise RuntimeError("%r has already been triggered" % (self,))
        self._ok, self._value = False, exception
        self.env._schedule(self, NORMAL, 0.0)
        return self

    def __or__(self, other):
        return AnyOf(self.env, [self, other])

    def __and__(self, other):
        return AllOf(self.env, [self, other])


class Timeout(Event):
    def __init__(self, env, delay, value=None):
        if delay < 0:
            raise ValueError("Negative delay %s" % delay)
        super().__init__(env)
        self._value = value
        env._schedule(self, NORMAL, delay)


class _Condition(Event):
    """``AnyOf`` / ``AllOf``: triggers with a dict {event: value} of the events processed so far."""

    def __init__(self, env, events, need_all):
        super().__init__(env)
        self._events = list(events)
        self._need_all = need_all
        self._count = 0
        if not self._events:
            self.succeed({})
            return
        for e in self._events:
            if e.callbacks is None:
                self._check(e)
            else:
                e.callbacks.append(self._check)

    def _check(self, event):
        if self._value is not _PENDING:
            return
        self._count += 1
        if not event._ok:
            self.fail(event._value)
        elif self._count == len(self._events) or not self._need_all:
            self.succeed({e: e._value for e in self._events if e.callbacks is None or e is event})


class AnyOf(_Condition):
    def __init__(self, env, events):
        super().__init__(env, events, False)


class AllOf(_Condition):
    def __init__(self, env, events):
        super().__init__(env, events, True)


class Process(Event):
    """Runs a generator: every yielded event suspends it until that event is processed; the process is
    itself an event that triggers with the generator's return value."""

    def __init__(self, env, generator):
        if not hasattr(generator, "send"):
            raise ValueError("%r is not a generator" % (generator,))
        super().__init__(env)
        self._generator = generator
        start = Event(env)
        start._value = None
        start.callbacks.append(self._resume)
        env._schedule(start, URGENT, 0.0)
        self._target = start

    @property
    def is_alive(self):
        return self._value is _PENDING

    def _resume(self, event):
        while True:
            try:
                if event._ok:
                    nxt = self._generator.send(event._value)
                else:
                    nxt = self._generator.throw(event._value)
            except StopIteration as stop:
                self._value = getattr(stop, "value", None)
                self.env._schedule(self, NORMAL, 0.0)
                return
            except BaseException as exc:
                self._ok, self._value = False, exc
                self.env._schedule(self, NORMAL, 0.0)
                raise
            if not isinstance(nxt, Event):
                raise RuntimeError("process yielded %r, which is not an event" % (nxt,))
            if nxt.callbacks is not None:       # not yet processed: wait for it
                nxt.callbacks.append(self._resume)
                self._target = nxt
                return
            event = nxt                         # already processed: continue right away


class Environment:
    def __init__(self, initial_time=0.0):
        self._now = initial_time
        self._queue = []
        self._eid = itertools.count()

    @property
    def now(self):
        return self._now

    def _schedule(self, event, priority, delay):
        heapq.heappush(self._queue, (self._now + delay, priority, next(self._eid), event))

    def event(self):
        return Event(self)

    def timeout(self, delay, value=None):
        return Timeout(self, delay, value)

    def process(self, generator):
        return Process(self, generator)

    def step(self):
        self._now, _, _, event = heapq.heappop(self._queue)
        callbacks, event.callbacks = event.callbacks, None
        for cb in callbacks:
            cb(event)

    def run(self, until=None):
        if until is not None and not isinstance(until, Event):
            at = float(until)
            if at <= self._now:
                raise ValueError("until (%s) must be greater than the current simulation time" % at)
            stop = Event(self)
            stop._value = None
            heapq.heappush(self._queue, (at, URGENT, next(self._eid), stop))
            until = stop
        if until is not None and until.callbacks is None:
            return until._value
        done = []
        if until is not None:
            until.callbacks.append(done.append)
        while self._queue and not done:
            self.step()
        if until is not None and not done:
            raise RuntimeError("no scheduled events left but the until-event was not triggered")
        return until._value if until is not None else None


class SimulationManager:
    """``gymwipe.simtools.SimulationManager`` (``simtools.py:18-128``) on the engine above."""

    def __init__(self):
        self._env = None

    @property
    def env(self):
        if self._env is None:
            self.init()
        return self._env

    @property
    def now(self):
        return self.env.now

    def init(self):
        """Creates a fresh environment (time 0, nothing scheduled)."""
        self._env = Environment()

    def process(self, generator):
        return self.env.process(generator)

    def event(self):
        return self.env.event()

    def timeout(self, duration, value=None):
        return self.env.timeout(duration, value)

    def timeoutUntil(self, triggerTime, value=None):
        """A timeout that fires at ``triggerTime`` (immediately if that lies in the past), ``simtools.py:103-116``."""
        now = self.now
        return self.env.timeout(triggerTime - now if triggerTime > now else 0, value)

    def nextTimeSlot(self, timeSlotLength):
        """A timeout until the next multiple of ``timeSlotLength`` (a full slot if on the grid), ``simtools.py:44-53``."""
        return self.env.timeout(timeSlotLength - self.now % timeSlotLength)

    def triggerAfterTimeout(self, event, timeout, value=None):
        def trigger():
            yield self.env.timeout(timeout)
            event.succeed(value)
        self.process(trigger())

    def runSimulation(self, until):
        """``until``: a duration (simulated seconds from now) or an event, ``simtools.py:77-88``."""
        if isinstance(until, Event):
            return self.env.run(until=until)
        return self.env.run(until=self.now + until)


SimMan = SimulationManager()


class Notifier:
    """
    ``gymwipe.simtools.Notifier`` (``simtools.py:232-432``): ``trigger(value)`` first runs the subscribed
    callbacks by descending priority value, then hands the value to the subscribed process factories and to every
    process waiting on :attr:`event`.

    Process subscriptions: ``blocking = False`` starts a new process instance for every trigger;
    ``blocking = True`` runs one instance at a time -- triggers that arrive meanwhile are dropped, or, with
    ``queued = True``, queued and processed one after the other.
    """

    def __init__(self, name="", owner=None):
        self._name = name
        self._owner = owner
        self._callbacks = {}            # priority -> list (subscription order)
        self._processes = []            # [factory, blocking, queued, running, queue]
        self._event = None

    def __repr__(self):
        return "Notifier('%s')" % self._name

    @property
    def name(self):
        return self._name

    def subscribeCallback(self, callback, priority=0, additionalArgs=None):
        """Calls ``callback(value, *additionalArgs)`` at every trigger; callbacks with a higher priority value
        run first (``simtools.py:310-318``)."""
        self._callbacks.setdefault(priority, []).append((callback, tuple(additionalArgs or ())))

    def unsubscribeCallback(self, callback):
        for entries in self._callbacks.values():
            entries[:] = [e for e in entries if e[0] is not callback and e[0] != callback]

    def subscribeProcess(self, process, blocking=True, queued=False):
        """``process(value)`` must return a generator; see the class docstring for the policies."""
        self._processes.append([process, blocking, queued, False, deque()])

    def _start(self, entry, value):
        factory, blocking, queued = entry[0], entry[1], entry[2]
        if not blocking:
            SimMan.process(factory(value))
            return
        entry[3] = True

        def finished(_event):
            # the instance has been processed: the next queued value, or idle again
            if queued and entry[4]:
                SimMan.process(factory(entry[4].popleft())).callbacks.append(finished)
            else:
                entry[3] = False
        SimMan.process(factory(value)).callbacks.append(finished)

    def trigger(self, value=None):
        for priority in sorted(self._callbacks, reverse=True):          # higher priority values first (simtools.py:310-318)
            for callback, extra in list(self._callbacks[priority]):
                callback(value, *extra)
        for entry in self._processes:
            if entry[1] and entry[3]:
                if entry[2]:
                    entry[4].append(value)
                continue
            self._start(entry, value)
        if self._event is not None:
            event, self._event = self._event, None
            event.succeed(value)

    @property
    def event(self):
        """An event that succeeds (with the value) at the next trigger; shared until then."""
        if self._event is None:
            self._event = SimMan.event()
        return self._event


def ensureType(obj, validTypes, caller=None):
    """Raises ``TypeError`` unless ``obj`` is an instance of ``validTypes`` (``simtools.py:212-229``)."""
    if not isinstance(obj, validTypes):
        raise TypeError("{}: Got object of invalid type {}. Expected type(s): {}".format(caller, type(obj), validTypes))
