"""Minimal environment objects.  Environment capture (conda/pip/docker, pylzy/lzy/env/*) is out of scope
for the data plane; ``Provisioning.gpu_count`` is read only to place an op."""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Any, Dict, Optional


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
