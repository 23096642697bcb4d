"""ORACLE TEST INFRASTRUCTURE -- ``pytest_mock.mocker`` fixture exposing ``Mock`` (the
reference's tests do ``from pytest_mock import mocker`` and use ``mocker.Mock()``)."""
from unittest import mock

import pytest


class _Mocker:
    Mock = mock.Mock
    MagicMock = mock.MagicMock
    patch = mock.patch


@pytest.fixture
def mocker():
    return _Mocker()
