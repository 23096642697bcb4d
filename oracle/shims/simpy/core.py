"""
ORACLE TEST INFRASTRUCTURE -- not product code.

Minimal engine compatible with ``simpy==3.0.11`` (the version pinned by the
reference, ``/root/reference/Pipfile.lock:176-183``).  simpy is a third-party
dependency that is NOT vendored under ``/root/reference`` and is not installed
in this image, so its *published* scheduling algorithm is restated here:

* the event heap is keyed ``(time, priority, eid)`` with ``eid`` a global
  insertion counter; ``URGENT = 0`` (process initialisation, ``run(until=<number>)``),
  ``NORMAL = 1`` (everything else);
* ``step()`` pops the minimum, sets ``now``, detaches the callback list
  (``event.callbacks = None`` is what ``Event.processed`` tests) and calls the
  callbacks in list order.

Only the reference's call sites need it (``gymwipe/simtools.py:11-13,58,68,75,88,95,101``,
``gymwipe/networking/physical.py:13,267,275,607``,
``gymwipe/networking/simple_stack.py:11,362,406-412,470-471``,
``gymwipe/networking/messages.py:25,204,225``).  The engine is accepted as the
oracle engine because the reference's own event-order tests pass on it
(``tests/test_simtools.py``, ``tests/networking/*``, ``tests/envs/*``; see
``oracle/run_reference_tests.py``).
"""
from heapq import heappop, heappush
from itertools import count

from simpy.events import (NORMAL, URGENT, AllOf, AnyOf, Event, Process,
                          Timeout)

Infinity = float('inf')


class EmptySchedule(Exception):
    """Raised by :meth:`Environment.step` when no event is left."""


class StopSimulation(Exception):
    """Raised (through an event callback) to stop :meth:`Environment.run`."""

    @classmethod
    def callback(cls, event):
        if event.ok:
            raise cls(event.value)
        raise event.value


class Environment:
    """Execution environment: simulation clock plus the event heap."""

    def __init__(self, initial_time=0):
        self._now = initial_time
        self._queue = []
        self._eid = count()
        self._active_proc = None
        # oracle-only instrumentation: number of heap pops
        self.popped_events = 0

    @property
    def now(self):
        return self._now

    @property
    def active_process(self):
        return self._active_proc

    # factories -----------------------------------------------------------
    def process(self, generator):
        return Process(self, generator)

    def timeout(self, delay, value=None):
        return Timeout(self, delay, value)

    def event(self):
        return Event(self)

    def all_of(self, events):
        return AllOf(self, events)

    def any_of(self, events):
        return AnyOf(self, events)

    # scheduling ----------------------------------------------------------
    def schedule(self, event, priority=NORMAL, delay=0):
        heappush(self._queue, (self._now + delay, priority, next(self._eid), event))

    def peek(self):
        try:
            return self._queue[0][0]
        except IndexError:
            return Infinity

    def step(self):
        try:
            self._now, _, _, event = heappop(self._queue)
        except IndexError:
            raise EmptySchedule()
        self.popped_events += 1

        callbacks, event.callbacks = event.callbacks, None
        for callback in callbacks:
            callback(event)

        if not event._ok and not hasattr(event, '_defused'):
            exc = type(event._value)(*event._value.args)
            exc.__cause__ = event._value
            raise exc

    def run(self, until=None):
        if until is not None:
            if not isinstance(until, Event):
                at = float(until)
                if at <= self.now:
                    raise ValueError('until(=%s) should be > the current '
                                     'simulation time.' % at)
                until = Event(self)
                until._ok = True
                until._value = None
                self.schedule(until, URGENT, at - self.now)
            elif until.callbacks is None:
                # already processed
                return until.value
            until.callbacks.append(StopSimulation.callback)

        try:
            while True:
                self.step()
        except StopSimulation as exc:
            return exc.args[0]
        except EmptySchedule:
            if until is not None:
                assert not until.triggered
                raise RuntimeError('No scheduled events left but "until" '
                                   'event was not triggered: %s' % until)
