"""
Declarative counterparts of ``gymwipe/networking/simple_stack.py``: the slotted-time PHY, the
contention-free MAC and the RRM MAC as scenario DESCRIPTORS.  Their behaviour (``simple_stack.py:32-561``)
is implemented by the CUDA step kernel (``gymwipe_b200/csrc/gw_core.cuh``); these objects carry the
parameters the reference keeps as attributes, and -- being ``Module`` s with the reference's ports
(``"mac"`` on the PHY, ``"phy"`` / ``"network"`` on the MACs) -- the wiring that
``gymwipe_b200.scenario.compile_stack`` traces into the scenario table (SURVEY.md section 8b:
"declarative scenario descriptors compiled into SoA tables").
"""
from gymwipe_b200.networking.construction import Module
from gymwipe_b200.networking.physical import BpskMcs, temperatureToNoisePowerDensity

TIME_SLOT_LENGTH = 1e-6
"""float: length of a time slot in seconds (``simple_stack.py:27``)"""


class SimplePhy(Module):
    """``simple_stack.py:32-286``."""

    NOISE_POWER_DENSITY = temperatureToNoisePowerDensity(20.0)

    def __init__(self, name, device, frequencyBand):
        super().__init__(name, owner=device)
        self.device = device
        self.frequencyBand = frequencyBand
        self._addPort("mac")                            # simple_stack.py:66
        self._thermalNoisePower = self.NOISE_POWER_DENSITY * frequencyBand.spec.bandwidth * 1000

    def __repr__(self):
        return "%r.SimplePhy('%s')" % (self.device, self.name)


class SimpleMac(Module):
    """``simple_stack.py:289-484``: queue of 100 packets, 0 dBm, BPSK 3/4."""

    rrmAddr = bytes(6)
    QUEUE_LENGTH = 100

    def __init__(self, name, device, frequencyBandSpec, addr):
        super().__init__(name, owner=device)
        self.device = device
        self.addr = addr
        self._addPort("phy")                            # simple_stack.py:354-355
        self._addPort("network")
        self._mcs = BpskMcs(frequencyBandSpec)
        self._transmissionPower = 0.0

    @staticmethod
    def macAddress(index):
        """Address number ``index`` of ``SimpleMac.newMacAddress`` (``simple_stack.py:376-384``)."""
        addr = bytearray(6)
        addr[5] = index & 0xFF
        addr[4] = (index >> 8) & 0xFF
        return bytes(addr)


class SimpleRrmMac(Module):
    """``simple_stack.py:486-561``: announcements at 0 dBm, one guard slot after each assignment."""

    def __init__(self, name, device, frequencyBandSpec):
        super().__init__(name, owner=device)
        self.device = device
        self.addr = bytes(6)
        self._addPort("phy")                            # simple_stack.py:517-518
        self._addPort("network")
        self._announcementMcs = BpskMcs(frequencyBandSpec)
        self._transmissionPower = 0.0
