"""Stub of nptyping: subscriptable annotation placeholders only."""


class _Sub:
    def __class_getitem__(cls, item):
        return cls

    def __getitem__(self, item):
        return self


class NDArray(_Sub):
    pass


class Shape(_Sub):
    pass


class Bool:
    pass


class Float32:
    pass


class Float64:
    pass


class UInt32:
    pass


class UInt64:
    pass


Float = Float64
