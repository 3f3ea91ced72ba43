"""``lzy.api.v1.utils.pickle`` by name: the request codec lives in :mod:`lzy_b200.utils.argv_codec`."""
from ....utils.argv_codec import decode as unpickle, encode as pickle

__all__ = ["pickle", "unpickle"]
