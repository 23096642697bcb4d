"""ORACLE TEST INFRASTRUCTURE -- gym.error names."""


class Error(Exception):
    pass


class UnregisteredEnv(Error):
    pass
