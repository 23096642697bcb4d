"""
Host-side compatibility layer for ``gymwipe/networking/construction.py``: ``Gate``, ``Port``,
``GateListener``, ``Module`` and ``CompoundModule`` with the reference's names and dispatch semantics
(OMNeT++-style plumbing).

In the reference this plumbing carries every packet of the simulation.  Here it is the vocabulary in
which a user WIRES a network stack -- the descriptor classes of ``gymwipe_b200.networking.simple_stack``
are ``Module`` s with the reference's ports -- and ``gymwipe_b200.scenario.compile_stack`` traces the
wiring into the scenario table the CUDA step kernel runs; messages sent through gates on the host are
dispatched exactly like the reference's (callbacks immediately, generator listeners as processes of
``gymwipe_b200.simtools.SimMan`` with the blocking / queued policies), so user-written modules and the
reference's own construction tests run unchanged.
"""
import inspect
from functools import wraps

from gymwipe_b200.simtools import Notifier, ensureType


def _prefix(owner):
    return "" if owner is None else "%r." % (owner,)


class Gate:
    """
    ``construction.py:20-111``: an object passed to :meth:`send` triggers :attr:`nReceives` and is thereby
    forwarded to every gate this gate was connected to with :meth:`connectTo`.
    """

    def __init__(self, name, owner=None):
        self.name = name
        self._owner = owner
        self.nReceives = Notifier('Receives', self)
        self.nConnectsTo = Notifier('Connects to', self)
        self.connections = []           # gates this gate forwards to (read by scenario.compile_stack)

    def __repr__(self):
        return "{}Gate('{}')".format(_prefix(self._owner), self.name)

    def connectTo(self, gate):
        self.nReceives.subscribeCallback(gate.send)
        self.connections.append(gate)
        self.nConnectsTo.trigger(gate)

    def send(self, object):
        self.nReceives.trigger(object)


class Port:
    """``construction.py:114-219``: an input and an output gate; bidirectional (proxy) connections."""

    def __init__(self, name, owner=None):
        self.name = name
        self._owner = owner
        self.input = Gate("in", owner=self)
        self.output = Gate("out", owner=self)

    def __repr__(self):
        return "{}Port('{}')".format(_prefix(self._owner), self.name)

    def biConnectWith(self, port):
        """my output -> its input, its output -> my input"""
        self.output.connectTo(port.input)
        port.output.connectTo(self.input)

    def biConnectProxy(self, port):
        """my output -> its output, its input -> my input (this port becomes the inner side of ``port``)"""
        self.output.connectTo(port.output)
        port.input.connectTo(self.input)

    @property
    def nReceives(self):
        return self.input.nReceives


class GateListener:
    """
    ``construction.py:221-342``: decorator factory -- the decorated method is called (or, if it is a
    generator function, run as a ``SimMan`` process) whenever ``self.gates[gateName]`` receives an object.
    ``blocking`` / ``queued`` select the process policy of ``Notifier.subscribeProcess``; the class'
    constructor has to be decorated with :meth:`setup`.
    """

    def __init__(self, gateName, validTypes=None, blocking=True, queued=False):
        self._gateName = gateName
        self._validTypes = validTypes
        self._blocking = blocking
        self._queued = queued

    def __call__(self, method):
        typecheck = self._validTypes is not None
        is_generator = inspect.isgeneratorfunction(method)
        listener = self

        def initializer(instance):
            def call_adapter(obj):
                if typecheck:
                    ensureType(obj, listener._validTypes, instance)
                return method(instance, obj)
            notifier = instance.gates[listener._gateName].nReceives
            if is_generator:
                notifier.subscribeProcess(call_adapter, listener._blocking, listener._queued)
            else:
                notifier.subscribeCallback(call_adapter)
        initializer.callAtConstruction = True
        initializer.__doc__ = "GateListener on gate `%s` (%s)." % (self._gateName, "process" if is_generator else "callback")
        return initializer

    @staticmethod
    def setup(function):
        """Decorator for the constructor of a ``Module`` subclass that uses ``GateListener``."""
        @wraps(function)
        def wrapper(self, *args, **kwargs):
            result = function(self, *args, **kwargs)
            for name in dir(self):
                if name.startswith("__"):
                    continue
                member = getattr(self, name)
                if getattr(member, "callAtConstruction", False):
                    member()
            return result
        return wrapper


class Module:
    """``construction.py:344-411``: a component with named ports (``ports``) and gates (``gates``)."""

    def __init__(self, name, owner=None):
        self.name = name
        self._owner = owner
        self.ports = {}
        self.gates = {}

    def __repr__(self):
        return "{}{}('{}')".format(_prefix(self._owner), self.__class__.__name__, self.name)

    def _addPort(self, name):
        if name in self.ports:
            raise ValueError("A port indexed by '{}' already exists.".format(name))
        port = Port(name, owner=self)
        self.ports[name] = port
        self.gates[name + "In"] = port.input
        self.gates[name + "Out"] = port.output

    def _addGate(self, name):
        if name in self.gates:
            raise ValueError("A gate indexed by '{}' already exists.".format(name))
        self.gates[name] = Gate(name, owner=self)


class CompoundModule(Module):
    """``construction.py:413-451``: a module made of submodules."""

    def __init__(self, name, owner=None):
        super().__init__(name, owner)
        self.submodules = {}

    def _addSubmodule(self, name, module):
        if name in self.submodules:
            raise ValueError("A submodule named '{}' already exists.".format(name))
        self.submodules[name] = module
