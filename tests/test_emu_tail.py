"""Tail primitives (GpbTailOp program ops) on the CPU fiber emulator against numpy."""
import pytest

import emu_util
import ops_cases as oc


@pytest.fixture(scope='module', autouse=True)
def emu():
    emu_util.attach()
    yield
    emu_util.detach()


def test_tail_primitives():
    oc.check_tail_primitives()
