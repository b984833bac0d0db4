"""pycsdr.types — Format and AgcProfile as the reference uses them:
Format members are passed to Buffer()/module constructors (csdr/chain/__init__.py:23,
csdr/chain/selector.py:33,121); AgcProfile is built by value, AgcProfile("Fast") (owrx/dsp.py:619)."""
from enum import Enum


class Format(Enum):
    CHAR = ("char", 1)
    SHORT = ("short", 2)
    FLOAT = ("float", 4)
    COMPLEX_FLOAT = ("complex_float", 8)
    COMPLEX_SHORT = ("complex_short", 4)
    COMPLEX_CHAR = ("complex_char", 2)

    @property
    def size(self):
        return self.value[1]


class AgcProfile(Enum):
    SLOW = "Slow"
    MID = "Mid"
    FAST = "Fast"
    LAGGY = "Laggy"
