"""
Declarative descriptor (a parameter holder, by design: SURVEY.md section 8b) with the names of ``gymwipe/control/inverted_pendulum.py``: the PID controller's parameters.  Its law
(``:45-69``: ``PID = kp*e + ki*(e + last_e) + kd*(e - last_e)`` on ``e = |angle|`` in degrees,
``+PID`` for negative and ``-PID`` for positive angles, one command every 10 ms) is evaluated by
the step kernel at the controller's traffic ticks.
"""


class InvertedPendulumPidController:
    def __init__(self, kp=1.0, ki=0.0, kd=0.0, controlInterval=0.01, payloadBytes=8):
        self.kp, self.ki, self.kd = kp, ki, kd
        self.controlInterval = controlInterval
        self.payloadBytes = payloadBytes
