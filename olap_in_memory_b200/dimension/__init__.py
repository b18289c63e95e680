from .catch_all import CatchAll
from .generic import GenericDimension
from .time import TimeDimension
from .timeslot import TimeSlot



class DimensionFactory:
    """/root/reference/src/dimension/factory.js:5-15"""

    @staticmethod
    def deserialize(buffer):
        from ..serialization import fromBuffer

        data = fromBuffer(buffer)
        if data.get("start"):
            return TimeDimension.deserialize(buffer)
        return GenericDimension.deserialize(buffer)


__all__ = ["CatchAll", "DimensionFactory", "GenericDimension", "TimeDimension", "TimeSlot"]
