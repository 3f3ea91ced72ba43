"""Two user serializers living in an importable module (a registry refuses serializers from ``__main__``): used by
tests/golden/registry_battery.py on both sides.  Duck-typed against the little that either registry asks of them."""


class Celsius:
    def __init__(self, degrees):
        self.degrees = degrees

    def __eq__(self, other):
        return isinstance(other, Celsius) and other.degrees == self.degrees


class _Base:
    def available(self):
        return True

    def stable(self):
        return True

    def requirements(self):
        return {}

    def meta(self):
        return {}

    def handles(self, typ):  # this package's registry asks this; serialzy's asks supported_types()
        supported = self.supported_types()
        return typ == supported if isinstance(supported, type) else bool(supported(typ))

    def serialize(self, obj, dest):
        self._serialize(obj, dest)

    def deserialize(self, source, typ=None):
        return self._deserialize(source, typ)


class CelsiusSerializer(_Base):
    def supported_types(self):
        return Celsius

    def data_format(self):
        return "battery_celsius"

    def _serialize(self, obj, dest):
        dest.write(repr(float(obj.degrees)).encode())

    def _deserialize(self, source, typ=None, user_type=None):
        return Celsius(float(source.read().decode()))


class EvenIntSerializer(_Base):
    def supported_types(self):
        return lambda t: False  # a predicate serializer: never chosen by type, only named by data format

    def data_format(self):
        return "battery_even_int"

    def _serialize(self, obj, dest):
        dest.write(str(obj).encode())

    def _deserialize(self, source, typ=None, user_type=None):
        return int(source.read().decode())
