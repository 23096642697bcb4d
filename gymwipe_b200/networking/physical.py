"""
Mirror of ``gymwipe/networking/physical.py``: dB/mW helpers, ``Mcs`` / ``BpskMcs``,
``Transmission``, ``FrequencyBandSpec`` / ``FrequencyBand``.

The scalar helpers (``physical.py:25-98``) are host-side conveniences with the reference's
exact expressions (they are configuration arithmetic, not the hot path).  The hot-path
arithmetic -- BER per SINR segment -- runs on the GPU: ``BpskMcs.calculateBitErrorRate`` on
tensors calls kernel K2 through the C ABI (``gw_ber_bpsk``).
"""
from fractions import Fraction
from math import e, log10, pi, sqrt

import torch

from gymwipe_b200 import _native as N


def calculateEbToN0Ratio(signalPower, noisePower, bitRate, returnDb=False):
    """``physical.py:25-42``: Eb/N0 from powers in dBm and the bit rate."""
    ratio_db = signalPower - noisePower - 10 * log10(bitRate)
    if returnDb:
        return ratio_db
    return 10 ** (ratio_db / 10)


sqrtOfTwoPi = sqrt(2 * pi)


def approxQFunction(x):
    """``physical.py:46-58``: Karagiannidis/Lioumpas approximation of the Gaussian Q function."""
    assert x >= 0
    return (1 - e ** (-1.4 * x)) * e ** (-(x ** 2 / 2)) / (1.135 * sqrtOfTwoPi * x)


def temperatureToNoisePowerDensity(temperature):
    """``physical.py:60-71``."""
    return 1.38e-23 * (temperature + 273.15)


def wattsToDbm(watts):
    return 10 * log10(watts) + 30


def milliwattsToDbm(milliwatts):
    return 10 * log10(milliwatts)


def dbmToMilliwatts(milliwatts):
    return 10 ** (milliwatts / 10)


class FrequencyBandSpec:
    """``physical.py:293-306``."""

    def __init__(self, frequency=2.4e9, bandwidth=22e6):
        self.frequency = frequency
        self.bandwidth = bandwidth


class Mcs:
    """``physical.py:100-185``: modulation and coding scheme."""

    def __init__(self, frequencyBandSpec, codeRate):
        self.frequencyBandSpec = frequencyBandSpec
        self.codeRate = codeRate

    def maxCorrectableBer(self):
        """Varshamov-Gilbert bound, evaluated by the native library (``gw_max_correctable_ber``)."""
        return float(N.lib().gw_max_correctable_ber(self.codeRate.numerator, self.codeRate.denominator))


class BpskMcs(Mcs):
    """``physical.py:187-212``: BPSK, 133.33333 kb/s on air, 100 kb/s of data at rate 3/4."""

    def __init__(self, frequencyBandSpec, codeRate=Fraction(3, 4)):
        super().__init__(frequencyBandSpec, codeRate)
        if codeRate != Fraction(3, 4):
            raise ValueError("the CUDA step kernel implements the reference's only MCS configuration "
                             "(BPSK, code rate 3/4)")
        self._bitRate = 133.33333e3
        self._dataRate = float(codeRate) * self._bitRate

    @property
    def bitRate(self):
        return self._bitRate

    @property
    def dataRate(self):
        return self._dataRate

    def calculateBitErrorRate(self, signalPower, noisePower):
        """
        BER for signal / noise powers in dBm (``physical.py:208-212``).  Tensors (CUDA,
        float64) are evaluated by kernel K2; Python floats use the same expression on the host.
        """
        if torch.is_tensor(signalPower):
            s = (10.0 ** (signalPower.double() / 10)).contiguous()
            n = (10.0 ** (noisePower.double() / 10)).contiguous()
            return ber_from_milliwatts(s, n)
        if signalPower <= noisePower:
            return 0.5
        ratio = calculateEbToN0Ratio(signalPower, noisePower, self._bitRate)
        return approxQFunction(sqrt(2 * ratio))


def ber_from_milliwatts(signal_mw, noise_mw):
    """Kernel K2 (``gw_ber_bpsk``): the exact evaluation ``SimplePhy._updateBitErrorRate`` does."""
    assert signal_mw.is_cuda and signal_mw.dtype == torch.float64
    out = torch.empty_like(signal_mw)
    stream = torch.cuda.current_stream(signal_mw.device).cuda_stream
    with torch.cuda.device(signal_mw.device):
        N.check(N.lib().gw_ber_bpsk(signal_mw.data_ptr(), noise_mw.data_ptr(), out.data_ptr(),
                                    signal_mw.numel(), stream))
    return out


class Transmission:
    """
    ``physical.py:214-290``.  Derived quantities of a packet on the air (durations, coded bit
    counts, stop time) with the reference's arithmetic; used for read-back and for tests.
    """

    def __init__(self, sender, power, headerBytes, payloadBytes, mcsHeader, mcsPayload, startTime):
        self.sender = sender
        self.power = power
        self.mcsHeader = mcsHeader
        self.mcsPayload = mcsPayload
        self.startTime = startTime
        self.headerDuration = headerBytes * 8 / mcsHeader.dataRate
        self.payloadDuration = payloadBytes * 8 / mcsPayload.dataRate
        self.duration = self.headerDuration + self.payloadDuration
        self.stopTime = startTime + self.duration
        self.headerBits = headerBytes * 8 * float(2 - mcsHeader.codeRate)
        self.payloadBits = payloadBytes * 8 * float(2 - mcsPayload.codeRate)

    def __repr__(self):
        return "Transmission(sender: {}, power: {} dBm, duration: {} s)".format(self.sender, self.power,
                                                                               self.duration)


class FrequencyBand:
    """``physical.py:530-655``: a wireless band with its attenuation model classes."""

    def __init__(self, modelClasses, frequency=2.4e9, bandwidth=22e6):
        from gymwipe_b200.networking.attenuation_models import FsplAttenuation
        if list(modelClasses) != [FsplAttenuation]:
            raise ValueError("the CUDA step kernel implements FsplAttenuation (the reference's only model)")
        self.modelClasses = list(modelClasses)
        self.spec = FrequencyBandSpec(frequency, bandwidth)
        self.devices = []       # filled by the scenario

    def __repr__(self):
        return "FrequencyBand(f={:.2E} Hz)".format(self.spec.frequency)
