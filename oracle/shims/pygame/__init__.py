"""ORACLE TEST INFRASTRUCTURE -- import stub: ``gymwipe/envs/__init__.py:4`` pulls in
``gymwipe/plants/sliding_pendulum.py:4-5`` which imports pygame (visualisation only)."""


class Surface:
    pass
