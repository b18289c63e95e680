from .catch_all import CatchAll
from .generic import GenericDimension
from .time import TimeDimension
from .timeslot import TimeSlot

__all__ = ["CatchAll", "GenericDimension", "TimeDimension", "TimeSlot"]
