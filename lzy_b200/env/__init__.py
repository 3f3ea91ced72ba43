"""Minimal environment objects.  Environment capture (conda/pip/docker, pylzy/lzy/env/*) is out of scope
for the data plane; ``Provisioning.gpu_count`` is read only to place an op.  What is kept is the
part an op can observe: environment variables declared on ``Lzy``, a workflow or an ``@op`` reach the op body
(pylzy/tests/scenarios/complex_graph/__init__.py:12-20), later levels overriding earlier ones
(pylzy/lzy/core/call.py:58-61), through the same ``with_env_vars`` / ``with_provisioning`` / ``with_env`` methods
(pylzy/lzy/env/mixin.py:18-75)."""
from __future__ import annotations

import contextlib
import os
import threading
from dataclasses import dataclass, field, replace
from typing import Any, Dict, Mapping, Optional


@dataclass(frozen=True)
class Provisioning:
    cpu_count: Optional[int] = None
    ram_size_gb: Optional[int] = None
    gpu_type: Optional[str] = None
    gpu_count: Optional[int] = None
    gpu_index: Optional[int] = None  # B200 runtime extension: pin an op to one GPU of the box


@dataclass(frozen=True)
class LzyEnvironment:
    env_vars: Dict[str, str] = field(default_factory=dict)
    provisioning: Provisioning = field(default_factory=Provisioning)
    namespace: Dict[str, Any] = field(default_factory=dict, repr=False, compare=False)

    def with_fields(self, **kwargs: Any) -> "LzyEnvironment":
        return replace(self, **kwargs)

    def combine(self, other: "LzyEnvironment") -> "LzyEnvironment":
        prov = other.provisioning if other.provisioning != Provisioning() else self.provisioning
        return LzyEnvironment({**self.env_vars, **other.env_vars}, prov, {**self.namespace, **other.namespace})

    def validate(self) -> None:
        pass

    def get_namespace(self) -> Dict[str, Any]:
        return self.namespace

    def get_provisioning(self) -> Provisioning:
        return self.provisioning

    def get_env_vars(self) -> Dict[str, str]:
        return self.env_vars


class WithEnvironmentMixin:
    """``with_env`` / ``with_env_vars`` / ``with_provisioning`` for anything that has an ``env`` and a
    ``with_fields`` (``Lzy``, ``LzyWorkflow``) or its own ``with_env`` (the ``@op`` wrapper).  Each call returns a
    new object; a level's declaration replaces that level's field, levels combine at call time."""
    env: LzyEnvironment

    def with_env(self, env: LzyEnvironment):
        return self.with_fields(env=env)  # type: ignore[attr-defined]

    def with_env_vars(self, *args: Mapping[str, str], **kwargs: str):
        if args and kwargs or len(args) > 1:
            raise NotImplementedError("wrong argument types")
        env_vars = dict(args[0]) if args else dict(kwargs)
        for k, v in env_vars.items():
            if not isinstance(k, str) or not isinstance(v, str):
                raise TypeError("environment variables are str → str")
        return self.with_env(self.env.with_fields(env_vars=env_vars))

    def with_provisioning(self, *args: Provisioning, **kwargs: Any):
        if args and kwargs or len(args) > 1:
            raise NotImplementedError("wrong argument types")
        if args and not isinstance(args[0], Provisioning):
            raise TypeError("bad argument types")
        provisioning = args[0] if args else Provisioning(**kwargs)
        return self.with_env(self.env.with_fields(provisioning=provisioning))


def _applier(method: str, *args: Any, **kwargs: Any):
    """decorator form of a ``with_*`` method (pylzy/lzy/env/shortcuts.py): ``@env_vars(A="1")`` above ``@op``, or
    applied to a ``Lzy`` / workflow object"""
    def wrapper(subject):
        if not isinstance(subject, WithEnvironmentMixin):
            raise TypeError(f"decorator/function {method.replace('with_', '')} can be applied only to "
                            f"@op-decorated functions, Lzy and LzyWorkflow objects")
        return getattr(subject, method)(*args, **kwargs)
    return wrapper


def env_vars(*args: Mapping[str, str], **kwargs: str):
    return _applier("with_env_vars", *args, **kwargs)


def provisioning(*args: Provisioning, **kwargs: Any):
    return _applier("with_provisioning", *args, **kwargs)


_ENV_COND = threading.Condition()
_ENV_IN_USE: Dict[str, list] = {}  # variable -> [value, how many running ops set it, what it was before]


@contextlib.contextmanager
def applied_env(env_vars: Optional[Mapping[str, str]]):
    """Run a block with `env_vars` in ``os.environ`` (restored afterwards).  The reference gives every op its own
    process; here the ops of one worker share one, so a variable belongs to the ops that are running with it: ops
    that agree on a value share it, an op that wants another value for the same name waits until the current users
    are done, and ops that declare no variables never wait for anybody."""
    if not env_vars:
        yield
        return
    wanted = {k: str(v) for k, v in env_vars.items()}
    with _ENV_COND:
        _ENV_COND.wait_for(lambda: all(k not in _ENV_IN_USE or _ENV_IN_USE[k][0] == v for k, v in wanted.items()))
        for k, v in wanted.items():
            slot = _ENV_IN_USE.get(k)
            if slot is None:
                _ENV_IN_USE[k] = [v, 1, os.environ.get(k)]
                os.environ[k] = v
            else:
                slot[1] += 1
    try:
        yield
    finally:
        with _ENV_COND:
            for k in wanted:
                slot = _ENV_IN_USE[k]
                slot[1] -= 1
                if slot[1] == 0:
                    del _ENV_IN_USE[k]
                    if slot[2] is None:
                        os.environ.pop(k, None)
                    else:
                        os.environ[k] = slot[2]
            _ENV_COND.notify_all()
