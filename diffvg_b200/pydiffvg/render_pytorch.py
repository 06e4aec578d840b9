from enum import IntEnum


class OutputType(IntEnum):
    color = 1
    sdf = 2
