"""lzy_b200.utils.numa: staging buffers next to the GPU that reads them.  Without a GPU (or without sysfs) every helper
must degrade to "leave things where they are" instead of failing."""
import os

import pytest

from lzy_b200.utils import numa


def test_cpulist_forms():
    assert numa._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert numa._parse_cpulist("5") == {5}
    assert numa._parse_cpulist("") == set()


def test_unknown_gpu_changes_nothing():
    before = os.sched_getaffinity(0)
    assert numa.local_cpus(123) is None
    with numa.near(123) as moved:
        assert moved is False
        assert os.sched_getaffinity(0) == before
    assert numa.bind_process(123) is False
    assert os.sched_getaffinity(0) == before


def test_near_moves_and_restores(monkeypatch):
    before = os.sched_getaffinity(0)
    one = {min(before)}
    monkeypatch.setattr(numa, "local_cpus", lambda gpu: one)
    with numa.near(0) as moved:
        assert moved is True and os.sched_getaffinity(0) == one
    assert os.sched_getaffinity(0) == before


def test_near_restores_after_an_exception(monkeypatch):
    before = os.sched_getaffinity(0)
    monkeypatch.setattr(numa, "local_cpus", lambda gpu: {min(before)})
    with pytest.raises(RuntimeError):
        with numa.near(0):
            raise RuntimeError("allocation failed")
    assert os.sched_getaffinity(0) == before


def test_cpus_outside_the_cpuset_fall_back_to_the_sysfs_list(monkeypatch):
    """a container may not be allowed on the GPU's socket at all: the attempt fails and nothing moves"""
    before = os.sched_getaffinity(0)
    monkeypatch.setattr(numa, "local_cpus", lambda gpu: {100000})
    with numa.near(0) as moved:
        assert moved is False
    assert os.sched_getaffinity(0) == before
