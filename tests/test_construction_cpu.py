"""
CPU suite: the Gate / Port / Module / GateListener / Notifier compatibility layer.
The scenarios are the reference's own tests -- ``tests/test_simtools.py:16-121`` and
``tests/networking/test_construction.py:18-200`` -- run against ``gymwipe_b200.simtools`` /
``gymwipe_b200.networking.construction`` (``unittest.mock`` instead of ``pytest_mock``), plus the tracer
that compiles a wired stack into the scenario table the CUDA step kernel runs.
"""
from unittest import mock

import pytest

from gymwipe_b200.networking.construction import CompoundModule, Gate, GateListener, Module, Port
from gymwipe_b200.simtools import Notifier, SimMan


@pytest.fixture
def simman():
    SimMan.init()
    yield SimMan


def test_notifier_callback(simman):                    # tests/test_simtools.py:16-43
    n = Notifier('myNotifier')
    value = "test1"
    history, callbacks = [], []
    for i in range(3, 0, -1):
        def callback(value, i=i):
            history.append((i, value))
        callbacks.append(callback)
    for priority, c in enumerate(callbacks):
        n.subscribeCallback(c, priority)
    n.trigger(value)
    assert history == [(i, value) for i in range(1, 4)]
    del history[:]
    for c in callbacks:
        n.unsubscribeCallback(c)
    n.trigger(value)
    assert history == []


def _logging_process(timeout):
    def proc(value):
        proc.instanceCounter += 1
        proc.value = value
        yield SimMan.timeout(timeout)
        proc.instanceCounter -= 1
    proc.value = None
    proc.instanceCounter = 0
    return proc


def test_notifier_process_policies(simman):            # tests/test_simtools.py:60-121
    n = Notifier("notifier")
    p1, p2, p3 = [_logging_process(10) for _ in range(3)]
    n.subscribeProcess(p1, blocking=False)
    n.subscribeProcess(p2, blocking=True, queued=False)
    n.subscribeProcess(p3, blocking=True, queued=True)

    def main():
        for i in range(1, 3):
            n.trigger("msg" + str(i))
            yield SimMan.timeout(1)
    SimMan.process(main())
    SimMan.runSimulation(4)
    assert p1.instanceCounter == 2 and p1.value == "msg2"
    assert p2.instanceCounter == 1 and p2.value == "msg1"
    assert p3.instanceCounter == 1 and p3.value == "msg1"
    SimMan.runSimulation(11)
    assert p1.instanceCounter == 0
    assert p2.instanceCounter == 0 and p2.value == "msg1"
    assert p3.instanceCounter == 1 and p3.value == "msg2"
    n.trigger("msg3")
    SimMan.runSimulation(1)
    assert p2.instanceCounter == 1 and p2.value == "msg3"
    SimMan.runSimulation(25)
    assert p1.instanceCounter == p2.instanceCounter == p3.instanceCounter == 0
    assert p1.value == p2.value == p3.value == "msg3"


def test_simman_time_primitives(simman):               # simtools.py:44-53, 103-116
    seen = []

    def proc():
        yield SimMan.timeout(2.5e-6)
        yield SimMan.nextTimeSlot(1e-6)
        seen.append(SimMan.now)
        yield SimMan.nextTimeSlot(1e-6)                # on the grid: a full slot
        seen.append(SimMan.now)
        yield SimMan.timeoutUntil(1.0)
        seen.append(SimMan.now)
        yield SimMan.timeoutUntil(0.5)                 # in the past: fires immediately
        seen.append(SimMan.now)
        a, b = SimMan.timeout(3, "a"), SimMan.timeout(1, "b")
        got = yield a | b
        seen.append((SimMan.now, list(got.values())))
    SimMan.process(proc())
    SimMan.runSimulation(10)
    assert seen[0] == pytest.approx(3e-6) and seen[1] == pytest.approx(4e-6)
    assert seen[2] == 1.0 and seen[3] == 1.0 and seen[4] == (2.0, ["b"])
    with pytest.raises(ValueError):
        SimMan.env.run(until=SimMan.now)


def test_ports():                                       # tests/networking/test_construction.py:18-40
    p1_receive, p2_receive = mock.Mock(), mock.Mock()
    p1 = Port("1")
    p1.input.nReceives.subscribeCallback(p1_receive)
    p2 = Port("2")
    p2.input.nReceives.subscribeCallback(p2_receive)
    p1.output.connectTo(p2.input)
    p2.output.connectTo(p1.input)
    p1.output.send('test message 1')
    p2_receive.assert_called_with('test message 1')
    p2.output.send('test message 2')
    p1_receive.assert_called_with('test message 2')


def test_module_functions():                            # tests/networking/test_construction.py:42-71
    m = Module('module1')
    m._addPort('port1')
    m._addPort('port2')
    assert m.ports['port1'].name == 'port1' and m.ports['port2'].name == 'port2'
    assert m.gates['port1In'] is m.ports['port1'].input and m.gates['port1Out'] is m.ports['port1'].output
    m._addGate('gate1')
    assert m.gates['gate1'].name == 'gate1'
    with pytest.raises(ValueError):
        m._addPort('port1')
    with pytest.raises(ValueError):
        m._addGate('gate1')
    cm = CompoundModule('CompoundModule')
    sub2, sub3 = Module('submodule2'), Module('submodule3')
    cm._addSubmodule('sub1', m)
    cm._addSubmodule('sub2', sub2)
    cm._addSubmodule('sub3', sub3)
    assert cm.submodules == {'sub1': m, 'sub2': sub2, 'sub3': sub3}
    with pytest.raises(ValueError):
        cm._addSubmodule('sub1', m)
    assert repr(m.gates['gate1']) == "Module('module1').Gate('gate1')"


def test_module_simulation(simman):                     # tests/networking/test_construction.py:73-135
    class TestModule(Module):
        def __init__(self, name):
            super().__init__(name)
            self._addPort("a")
            self._addPort("b")
            self.msgReceivedCount = {"a": 0, "b": 0}
            self.msgVal = None
            SimMan.process(self.process("a", "b"))
            SimMan.process(self.process("b", "a"))

        def process(self, fromPort, toPort):
            while True:
                msg = yield self.ports[fromPort].nReceives.event
                self.msgVal = msg
                self.msgReceivedCount[fromPort] += 1
                msg += 1
                yield SimMan.env.timeout(1)
                if msg % 10 == 0:
                    self.ports[fromPort].output.send(msg)
                else:
                    self.ports[toPort].output.send(msg)

    m1, m2 = TestModule("1"), TestModule("2")
    m1.ports["b"].biConnectWith(m2.ports["b"])
    m2.ports["a"].biConnectWith(m1.ports["a"])
    checked = []

    def simulation():
        m1.gates["aIn"].send(1)
        yield SimMan.timeout(20)
        assert m1.msgVal == 19 and m2.msgVal == 20
        yield SimMan.timeout(20)
        for m in (m1, m2):
            for port in ("a", "b"):
                assert m.msgReceivedCount[port] == 10
        checked.append(True)
    SimMan.process(simulation())
    SimMan.runSimulation(50)
    assert checked == [True]


class MyModule(Module):
    @GateListener.setup
    def __init__(self, name):
        super().__init__(name)
        self._addPort("a")
        self._addPort("b")
        self.logs = [[] for _ in range(4)]

    @GateListener("aIn", queued=False)
    def aListener(self, message):
        self.logs[0].append(message)

    @GateListener("aIn", queued=True)
    def aListenerQueued(self, message):
        self.logs[1].append(message)

    @GateListener("bIn", queued=False)
    def bListener(self, message):
        self.logs[2].append(message)
        yield SimMan.timeout(10)

    @GateListener("bIn", queued=True)
    def bListenerQueued(self, message):
        self.logs[3].append(message)
        yield SimMan.timeout(10)


def test_gate_listener_method(simman):                  # tests/networking/test_construction.py:165-176
    modules = MyModule("Test1"), MyModule("Test2")
    for i in range(3):
        for module in modules:
            module.gates["aIn"].send("msg" + str(i))
            for j in range(2):
                assert module.logs[j] == ["msg" + str(n) for n in range(i + 1)]


def test_gate_listener_generator(simman):               # tests/networking/test_construction.py:178-200
    modules = MyModule("Test1"), MyModule("Test2")

    def main():
        for i in range(3):
            for module in modules:
                module.gates["bIn"].send("msg" + str(i))
                yield SimMan.timeout(1)
    SimMan.process(main())
    SimMan.runSimulation(40)
    for module in modules:
        assert module.logs[2] == ["msg0"]
        assert module.logs[3] == ["msg" + str(n) for n in range(3)]


def test_gate_listener_typecheck(simman):
    class Typed(Module):
        @GateListener.setup
        def __init__(self):
            super().__init__("typed")
            self._addGate("in")
            self.got = []

        @GateListener("in", int)
        def listener(self, obj):
            self.got.append(obj)
    t = Typed()
    t.gates["in"].send(3)
    assert t.got == [3]
    with pytest.raises(TypeError):
        t.gates["in"].send("three")


# ---------------------------------------------------------------------------------------------------------
# tracing wired stacks into the scenario table
# ---------------------------------------------------------------------------------------------------------

def _wired_band():
    """CounterTrafficEnv's devices (counter_traffic.py:124-133), wired by hand with the reference's plumbing --
    sender 1's PHY and MAC through a proxy port, as tests/networking/test_stack.py:134-158 does."""
    from gymwipe_b200.networking.attenuation_models import FsplAttenuation
    from gymwipe_b200.networking.devices import NetworkDevice, PhySenderDevice
    from gymwipe_b200.networking.physical import FrequencyBand
    from gymwipe_b200.networking.simple_stack import SimpleMac, SimplePhy, SimpleRrmMac
    band = FrequencyBand([FsplAttenuation])

    class Sender(NetworkDevice):
        def __init__(self, name, x, y, mult, index, proxy=False):
            super().__init__(name, x, y, band)
            self.packetMultiplicity = mult
            self.phy = SimplePhy("phy", self, band)
            self.mac = SimpleMac("mac", self, band.spec, SimpleMac.macAddress(index))
            if proxy:
                self.proxy = Port("proxy")
                self.phy.ports["mac"].biConnectProxy(self.proxy)
                self.proxy.biConnectWith(self.mac.ports["phy"])
            else:
                self.mac.ports["phy"].biConnectWith(self.phy.ports["mac"])

    class Rrm(NetworkDevice):
        def __init__(self):
            super().__init__("RRM", 0.0, 0.0, band)
            self.phy = SimplePhy("phy", self, band)
            self.mac = SimpleRrmMac("mac", self, band.spec)
            self.mac.ports["phy"].biConnectWith(self.phy.ports["mac"])

    rrm = Rrm()                                         # construction order does not matter: roles come from the wiring
    s1 = Sender("Sender 1", 0.0, 2.0, 1, 1, proxy=True)
    s2 = Sender("Sender 2", 0.0, -2.0, 3, 2)
    return band, (s1, s2, rrm), PhySenderDevice


def test_compile_stack_reproduces_the_default_scenario():
    from gymwipe_b200.scenario import compile_stack, config_from_dict, default_scenario_dict
    band, _, _ = _wired_band()
    sc = compile_stack([band])
    assert sc == default_scenario_dict()
    cfg = config_from_dict(sc, 8)
    assert cfg.band[0].n_devices == 3 and cfg.band[0].device[1].multiplicity == 3


def test_compile_stack_phy_only_sender_and_errors():
    from gymwipe_b200.networking.simple_stack import SimplePhy
    from gymwipe_b200.scenario import compile_stack
    band, (s1, s2, rrm), PhySenderDevice = _wired_band()
    PhySenderDevice("Jammer", 6.0, 0.0, band, sendInterval=0.05, initialDelay=0.003, power=40.0, payloadBytes=26)
    sc = compile_stack([band], assignment_duration_factor=10000)
    devs = sc["bands"][0]["devices"]
    assert [d["role"] for d in devs] == ["sender", "sender", "rrm", "jammer"]
    assert devs[3] == {"role": "jammer", "x": 6.0, "y": 0.0, "interval": 0.05, "delay": 0.003, "power": 40.0, "hdr": 13, "payload": 26}
    # a MAC whose phy port is not wired back to the PHY is not a working stack
    band2, (a, b, r), _ = _wired_band()
    b.mac.ports["phy"].output.connections.clear()
    with pytest.raises(ValueError, match="not connected back"):
        compile_stack([band2])
    # a device without a PHY on the band
    band3, (a, b, r), _ = _wired_band()
    del a.phy
    with pytest.raises(ValueError, match="exactly one SimplePhy"):
        compile_stack([band3])


@pytest.mark.gpu
def test_compiled_stack_runs_on_the_step_kernel():
    """The traced table drives the CUDA step kernel and passes the reference's known-answer test
    (tests/envs/test_counter_traffic.py:25-34)."""
    import gymwipe_b200
    from gymwipe_b200.scenario import compile_stack
    band, _, _ = _wired_band()
    env = gymwipe_b200.make('CounterTraffic-v0', num_envs=1, scenario=compile_stack([band]))
    env.reset()
    obs, reward, done, info = env.step({"device": 0, "duration": 3})
    assert obs - env.COUNTER_BOUND == 2 and reward == -2
    obs, reward, done, info = env.step({"device": 1, "duration": 12})
    assert obs - env.COUNTER_BOUND == 0 and reward == 2


def _wired_band_n(n_senders=4, n_phy_only=2):
    """A band beyond CounterTrafficEnv's template, wired by hand: n MAC senders on a circle, each addressing the
    next one, the RRM in the middle, PHY-only senders outside."""
    import math
    from gymwipe_b200.networking.attenuation_models import FsplAttenuation
    from gymwipe_b200.networking.devices import NetworkDevice, PhySenderDevice
    from gymwipe_b200.networking.physical import FrequencyBand
    from gymwipe_b200.networking.simple_stack import SimpleMac, SimplePhy, SimpleRrmMac
    band = FrequencyBand([FsplAttenuation])

    class Sender(NetworkDevice):
        def __init__(self, k):
            a = 2 * math.pi * k / n_senders
            super().__init__("Sender %d" % k, 2 * math.cos(a), 2 * math.sin(a), band)
            self.packetMultiplicity = 1 + k % 3
            self.phy = SimplePhy("phy", self, band)
            self.mac = SimpleMac("mac", self, band.spec, SimpleMac.macAddress(k + 1))
            self.mac.ports["phy"].biConnectWith(self.phy.ports["mac"])

    class Rrm(NetworkDevice):
        def __init__(self):
            super().__init__("RRM", 0.0, 0.0, band)
            self.phy = SimplePhy("phy", self, band)
            self.mac = SimpleRrmMac("mac", self, band.spec)
            self.mac.ports["phy"].biConnectWith(self.phy.ports["mac"])

    senders = [Sender(k) for k in range(n_senders)]
    for k, s in enumerate(senders):
        s.destination = senders[(k + 1) % n_senders]
    rrm = Rrm()
    for j in range(n_phy_only):
        PhySenderDevice("Interferer %d" % j, 6.0 + j, 0.0, band, sendInterval=0.02 + 0.005 * j, initialDelay=0.001 * j, power=10.0,
                        payloadBytes=40)
    return band, senders, rrm


def test_compile_stack_bands_beyond_the_template():
    from gymwipe_b200.envs import fits_step_kernel_template
    from gymwipe_b200.scenario import compile_stack
    band, senders, rrm = _wired_band_n(4, 2)
    sc = compile_stack([band])
    devs = sc["bands"][0]["devices"]
    assert [d["role"] for d in devs] == ["sender"] * 4 + ["rrm"] + ["jammer"] * 2
    assert [d["dest"] for d in devs[:4]] == [1, 2, 3, 0] and [d["mult"] for d in devs[:4]] == [1, 2, 3, 1]
    assert not fits_step_kernel_template(sc)
    del senders[2].destination
    with pytest.raises(ValueError, match="destination"):
        compile_stack([band])
    band9, _, _ = _wired_band_n(9, 0)
    with pytest.raises(ValueError, match="2..8 MAC senders"):
        compile_stack([band9])


@pytest.mark.gpu
def test_compiled_stack_beyond_the_template_runs_on_the_general_band_engine():
    """A hand-wired band of 4 senders + RRM + 2 PHY-only senders compiles to a table that `make` hands to the general
    band engine; its steps equal the oracle's."""
    import numpy as np
    import torch
    import gymwipe_b200
    import gw_oracle as O
    from gymwipe_b200.envs import GeneralBandEnv
    from gymwipe_b200.scenario import compile_stack
    band, _, _ = _wired_band_n(4, 2)
    sc = compile_stack([band])
    nenv, T = 64, 30
    rs = np.random.RandomState(12)
    dev = rs.randint(0, 4, size=(T, nenv)).astype(np.int32)
    dur = rs.randint(0, 20, size=(T, nenv)).astype(np.int32)
    o = O.run_batch(sc, dev, dur)
    env = gymwipe_b200.make('CounterTraffic-v0', num_envs=nenv, scenario=sc, strict=False)
    assert isinstance(env, GeneralBandEnv)
    env.reset()
    for t in range(T):
        obs, rew, done, _ = env.step({"device": torch.as_tensor(dev[t]).cuda(), "duration": torch.as_tensor(dur[t]).cuda()})
        assert (obs.cpu().numpy() == o["obs"][t, :, 0]).all() and (rew.cpu().numpy() == o["reward"][t, :, 0]).all()
        assert (env.now.cpu().numpy() == o["now"][t]).all()
    env.check()
    assert (env.delivered().cpu().numpy() == o["counts"][:, 0, 1:5]).all() and o["counts"][:, 0, 1:5].sum() > 0
