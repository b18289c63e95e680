"""Single-item placeholder dimension that Cube.addDimension inserts before a
drillDown (/root/reference/src/dimension/catch-all.js:3-69)."""
from __future__ import annotations

import numpy as np

from .abstract import AbstractDimension


class CatchAll(AbstractDimension):
    def __init__(self, id, childDimension=None):
        super().__init__(id, "all")
        self.childDimension = childDimension

    @property
    def attributes(self):
        raise NotImplementedError("Unsupported")

    def serialize(self):  # catch-all.js:16-18
        raise NotImplementedError("Unsupported")

    def getItems(self, _attribute=None):
        return ["_total"]

    def getEntries(self, _attribute=None, _language="en"):
        return [["_total", "Total"]]

    def drillUp(self, _newAttribute):
        return self

    def drillDown(self, newAttribute):
        if self.childDimension is not None:
            return self.childDimension.drillUp(newAttribute)
        raise ValueError("Must set child dimension.")

    def dice(self, attribute, items, _reorder=False):
        if attribute == self.rootAttribute and "_total" in items:
            return self
        raise NotImplementedError("Unsupported")

    def diceRange(self, _attribute, _start, _end):
        raise NotImplementedError("Unsupported")

    def getGroupIndexFromRootIndex(self, _attribute, _index):
        return 0

    def getGroupIndexFromRootIndexMap(self, _attribute):
        return np.zeros(1, dtype=np.int32)

    def intersect(self, otherDimension):
        return otherDimension

    def union(self, _otherDimension):
        return self
