"""Calendar dimension: a closed day range viewed at some periodicity.
Host-side mirror of /root/reference/src/dimension/time.js:5-249.

The device path only sees the root-index -> group-index map built by
getGroupIndexFromRootIndexMap (time.js:182-197): monotone non-decreasing
int32, e.g. 3652 days -> 120 months for 2010-01-01..2019-12-31."""
from __future__ import annotations

import numpy as np

from .abstract import AbstractDimension
from .timeslot import TimeSlot


class TimeDimension(AbstractDimension):
    def __init__(self, id, rootAttribute, start, end, label=None):
        super().__init__(id, rootAttribute, label)
        self._start = TimeSlot.fromDate(TimeSlot.fromValue(start).firstDate, "day")
        self._end = TimeSlot.fromDate(TimeSlot.fromValue(end).lastDate, "day")
        self._items = {}
        self._rootIdxToGroupIdx = {}
        self._drilled = {}  # attribute -> dimension; time dimensions are immutable

    @property
    def attributes(self):
        return [self._rootAttribute, *TimeSlot.upperSlots[self._rootAttribute]]

    @staticmethod
    def deserialize(buffer):  # time.js:28-37
        from ..serialization import fromBuffer

        data = fromBuffer(buffer)
        return TimeDimension(data["id"], data["rootAttribute"], data["start"], data["end"], data["label"])

    def serialize(self):  # time.js:39-47
        from ..serialization import toBuffer

        return toBuffer({
            "id": self.id,
            "label": self.label,
            "rootAttribute": self.rootAttribute,
            "start": self._start.value,
            "end": self._end.value,
        })

    def getItems(self, attribute=None):
        if self._start.value > self._end.value:
            return []
        attr = attribute or self._rootAttribute
        items = self._items.get(attr)
        if items is None:
            end = self._end.toParentPeriodicity(attr)
            period = self._start.toParentPeriodicity(attr)
            items = [period.value]
            while period.value < end.value:
                period = period.next()
                items.append(period.value)
            self._items[attr] = items
        return items

    def getEntries(self, attribute=None, language="en"):
        return [[item, TimeSlot.fromValue(item).humanizeValue(language)] for item in self.getItems(attribute)]

    def drillUp(self, newAttribute):
        if newAttribute == self.rootAttribute:
            return self
        return self._drill(newAttribute)

    def _drill(self, newAttribute):
        dim = self._drilled.get(newAttribute)
        if dim is None:
            dim = TimeDimension(self.id, newAttribute, self._start.value, self._end.value, self.label)
            self._drilled[newAttribute] = dim
        return dim

    def drillDown(self, newAttribute):
        if newAttribute == self.rootAttribute:
            return self
        if self._rootAttribute not in TimeSlot.upperSlots[newAttribute]:
            raise ValueError("Invalid periodicity.")
        return self._drill(newAttribute)

    def dice(self, attribute, items, reorder=False):
        if len(items) == 1:
            return self.diceRange(attribute, items[0], items[0])
        working = list(items) if reorder else sorted(items)
        last = TimeSlot.fromValue(items[0])
        if last.periodicity != attribute:
            raise ValueError("Unsupported: wrong periodicity")
        for value in working[1:]:
            current = TimeSlot.fromValue(value)
            if current.periodicity != attribute or current.value != last.next().value:
                raise ValueError("Unsupported: follow")
            last = current
        return self.diceRange(attribute, working[0], working[-1])

    def diceRange(self, attribute, start, end):
        if attribute == "all":
            return self
        if start:
            slot = TimeSlot.fromValue(start)
            if slot.periodicity != attribute:
                raise ValueError(f"{start} is not a valid slot of periodicity {attribute}")
            new_start = TimeSlot.fromDate(slot.firstDate, "day").value
        else:
            new_start = self._start.value
        if end:
            slot = TimeSlot.fromValue(end)
            if slot.periodicity != attribute:
                raise ValueError(f"{end} is not a valid slot of periodicity {attribute}")
            new_end = TimeSlot.fromDate(slot.lastDate, "day").value
        else:
            new_end = self._end.value
        if new_start <= self._start.value and self._end.value <= new_end:
            return self
        return TimeDimension(
            self.id,
            self._rootAttribute,
            max(new_start, self._start.value),
            min(new_end, self._end.value),
            self.label,
        )

    def getGroupIndexFromRootIndexMap(self, groupAttr):
        mapping = self._rootIdxToGroupIdx.get(groupAttr)
        if mapping is None:
            self._checkAttribute(groupAttr)
            group_idx = self.getItemsToIdx(groupAttr)
            mapping = np.fromiter(
                (
                    group_idx[TimeSlot.fromValue(item).toParentPeriodicity(groupAttr).value]
                    for item in self.getItems()
                ),
                dtype=np.int32,
                count=self.numItems,
            )
            self._rootIdxToGroupIdx[groupAttr] = mapping
        return mapping

    def getGroupIndexFromRootIndex(self, groupAttr, rootIdx):
        return int(self.getGroupIndexFromRootIndexMap(groupAttr)[rootIdx])

    def union(self, otherDimension):
        if self.id != otherDimension.id:
            raise ValueError("Not the same dimension")
        if otherDimension.rootAttribute in self.attributes:
            root = otherDimension._rootAttribute
        elif self.rootAttribute in otherDimension.attributes:
            root = self._rootAttribute
        else:
            raise ValueError("The dimensions are not compatible")
        start = min(self._start.value, otherDimension._start.value)
        end = max(self._end.value, otherDimension._end.value)
        return TimeDimension(self.id, root, start, end, self.label)

    def intersect(self, otherDimension):
        if self.id != otherDimension.id:
            raise ValueError("Not the same dimension")
        if otherDimension.rootAttribute in self.attributes:
            return otherDimension.diceRange("day", self._start.value, self._end.value)
        if self.rootAttribute in otherDimension.attributes:
            return self.diceRange("day", otherDimension._start.value, otherDimension._end.value)
        raise ValueError("The dimensions are not compatible")
