"""
Declarative mirror of ``gymwipe.networking`` (model semantics of FrequencyBand /
Transmission / SimplePhy / SimpleMac / SimpleRrmMac).  In the reference these classes ARE
the simulation (SimPy processes and callbacks); here they are scenario descriptors that
``gymwipe_b200.scenario`` compiles into the structure-of-arrays tables the CUDA step kernel
consumes, plus read-back views of the device state.
"""
from gymwipe_b200.networking.attenuation_models import FsplAttenuation
from gymwipe_b200.networking.devices import (Device, NetworkDevice, Position, SimpleNetworkDevice,
                                             SimpleRrmDevice)
from gymwipe_b200.networking.physical import (BpskMcs, FrequencyBand, FrequencyBandSpec, Mcs,
                                              Transmission, approxQFunction, calculateEbToN0Ratio,
                                              dbmToMilliwatts, milliwattsToDbm,
                                              temperatureToNoisePowerDensity, wattsToDbm)
from gymwipe_b200.networking.simple_stack import (TIME_SLOT_LENGTH, SimpleMac, SimplePhy,
                                                  SimpleRrmMac)

__all__ = ["FsplAttenuation", "Device", "NetworkDevice", "Position", "SimpleNetworkDevice",
           "SimpleRrmDevice", "BpskMcs", "FrequencyBand", "FrequencyBandSpec", "Mcs", "Transmission",
           "approxQFunction", "calculateEbToN0Ratio", "dbmToMilliwatts", "milliwattsToDbm",
           "temperatureToNoisePowerDensity", "wattsToDbm", "TIME_SLOT_LENGTH", "SimpleMac", "SimplePhy",
           "SimpleRrmMac"]
