"""
Mirror of ``gymwipe/devices/core.py`` and ``gymwipe/networking/devices.py``: positions and
network devices as scenario descriptors.
"""
from math import sqrt

from gymwipe_b200.networking.simple_stack import SimpleMac, SimplePhy, SimpleRrmMac


class Position:
    """``devices/core.py:15-98`` (default position of a device; per-env positions are tensors)."""

    def __init__(self, x, y, owner=None):
        self._x = float(x)
        self._y = float(y)
        self._owner = owner

    @property
    def x(self):
        return self._x

    @property
    def y(self):
        return self._y

    def __eq__(self, p):
        return p.x == self._x and p.y == self._y

    def distanceTo(self, p):
        return sqrt((self.x - p.x) ** 2 + (self.y - p.y) ** 2)

    def __repr__(self):
        return "Position({},{})".format(self.x, self.y)


class Device:
    """``devices/core.py:100-123``."""

    def __init__(self, name, xPos, yPos):
        self.name = name
        self._position = Position(xPos, yPos, self)

    @property
    def position(self):
        return self._position

    def __repr__(self):
        return "Device('{}')".format(self.name)


class NetworkDevice(Device):
    """``networking/devices.py:14-38``."""

    def __init__(self, name, xPos, yPos, frequencyBand):
        super().__init__(name, xPos, yPos)
        self.frequencyBand = frequencyBand
        frequencyBand.devices.append(self)


class SimpleNetworkDevice(NetworkDevice):
    """``networking/devices.py:40-111``: SimplePhy + SimpleMac."""

    _role = "sender"

    def __init__(self, name, xPos, yPos, frequencyBand, macIndex):
        super().__init__(name, xPos, yPos, frequencyBand)
        self.macAddr = SimpleMac.macAddress(macIndex)
        self._phy = SimplePhy("phy", self, frequencyBand)
        self._mac = SimpleMac("mac", self, frequencyBand.spec, self.macAddr)
        self._mac.ports["phy"].biConnectWith(self._phy.ports["mac"])            # devices.py:58


class SimpleRrmDevice(NetworkDevice):
    """``networking/devices.py:113-203``: SimplePhy + SimpleRrmMac + interpreter."""

    _role = "rrm"

    def __init__(self, name, xPos, yPos, frequencyBand, deviceIndexToMacDict, interpreter):
        super().__init__(name, xPos, yPos, frequencyBand)
        self.interpreter = interpreter
        self.deviceIndexToMacDict = deviceIndexToMacDict
        self.macToDeviceIndexDict = {mac: index for index, mac in deviceIndexToMacDict.items()}
        self._phy = SimplePhy("phy", self, frequencyBand)
        self._mac = SimpleRrmMac("mac", self, frequencyBand.spec)
        self._mac.ports["phy"].biConnectWith(self._phy.ports["mac"])            # devices.py:131

    @property
    def macAddr(self):
        return self._mac.addr


class PhySenderDevice(NetworkDevice):
    """PHY-only periodic sender ("jammer"), after ``tests/test_benchmark.py:20-50``."""

    _role = "jammer"

    def __init__(self, name, xPos, yPos, frequencyBand, sendInterval, initialDelay, power=0.0,
                 headerBytes=13, payloadBytes=26):
        super().__init__(name, xPos, yPos, frequencyBand)
        self.sendInterval = float(sendInterval)
        self.initialDelay = float(initialDelay)
        self.power = float(power)
        self.headerBytes = int(headerBytes)
        self.payloadBytes = int(payloadBytes)
        self._phy = SimplePhy("phy", self, frequencyBand)
