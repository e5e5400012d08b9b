"""``backpack.extensions`` stand-in: only the marker class the reference imports."""


class SumGradSquared:
    pass
