"""The part of ``lzy.serializer_registry`` that pylzy adds on top of serialzy's registry, exercised twice: by
tests/golden/make_golden.py on the *reference's own* ``LzySerializerRegistry`` (pylzy/lzy/serialization/registry.py:20-82,
loaded by file path in the authoring container over a stand-in base class, serialzy being absent) →
tests/golden/reference_registry_vectors.json, and by tests/test_api_core.py on lzy_b200.serialization.registry's:
user serializers are remembered by class and shipped as (module, class, priority) imports, built-ins are not;
``load_imports`` brings them up in a fresh registry; a serializer from ``__main__`` is refused; unregistering forgets;
values of a registered type nested inside a pickle go through their serializer (copyreg).

`ns["new_registry"]() -> registry`; ``SerializerImport`` of that side."""
import copyreg
import pickle

import cloudpickle
from typing import Any, Dict

import importlib.util
import sys
from pathlib import Path

_NAME = "lzyb_battery_serializers"  # an importable home for the user serializers (a registry refuses ``__main__``)
if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, Path(__file__).with_name("battery_serializers.py"))
    _module = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _module
    _spec.loader.exec_module(_module)
Celsius, CelsiusSerializer, EvenIntSerializer = (getattr(sys.modules[_NAME], n) for n in
                                                 ("Celsius", "CelsiusSerializer", "EvenIntSerializer"))


def run(ns: Dict[str, Any]) -> Dict[str, Any]:
    out: Dict[str, Any] = {}
    shown = lambda imports: sorted([i.module_name, i.class_name, i.priority] for i in imports)  # noqa: E731

    reg = ns["new_registry"]()
    out["fresh registry ships nothing"] = shown(reg.imports())
    reg.register_serializer(CelsiusSerializer(), 5)
    reg.register_serializer(EvenIntSerializer(), 9)
    out["after two registrations"] = shown(reg.imports())
    out["celsius is hooked into pickles"] = Celsius in copyreg.dispatch_table
    nested = pickle.loads(cloudpickle.dumps({"t": [Celsius(21.5)]}))  # cloudpickle: what the pickle fallback serializer uses
    out["nested value round trip"] = nested == {"t": [Celsius(21.5)]}

    other = ns["new_registry"]()
    other.load_imports(reg.imports())
    out["loaded elsewhere"] = shown(other.imports())
    out["found by data format there"] = type(other.find_serializer_by_data_format("battery_celsius")).__name__

    main_like = type("FromMain", (CelsiusSerializer,), {"__module__": "__main__"})
    try:
        reg.register_serializer(main_like(), 3)
        out["from __main__"] = "accepted"
    except ValueError as e:
        out["from __main__"] = str(e)

    reg.unregister_serializer(CelsiusSerializer())
    out["after unregistering celsius"] = shown(reg.imports())
    out["celsius hook removed"] = Celsius not in copyreg.dispatch_table
    # (the copyreg table is process-wide: the reference raises KeyError when a second registry unregisters a type whose
    # hook the first one already removed — this package's tolerates it; the battery does not go there)
    other.unregister_serializer(EvenIntSerializer())
    reg.unregister_serializer(EvenIntSerializer())
    out["what is left"] = [shown(reg.imports()), shown(other.imports())]
    return out
