"""Common dimension behaviour (host side).  Mirrors the members of
/root/reference/src/dimension/abstract.js:1-104 that the cube-transform path
uses: numItems, rootAttribute, getItems, getItemsToIdx, getRootIndexFromRootItem
and the group-index helpers.  Dimensions are O(items) string/graph objects; the
only thing that crosses the C ABI is the Int32 map they produce
(getGroupIndexFromRootIndexMap)."""
from __future__ import annotations


class AbstractDimension:
    def __init__(self, id, rootAttribute, label=None):
        self.id = id
        self._rootAttribute = rootAttribute
        self._label = label
        self._itemsToIdx = {}

    @property
    def numItems(self):
        return len(self.getItems())

    @property
    def rootAttribute(self):
        return self._rootAttribute

    @property
    def attributes(self):
        raise NotImplementedError("Override me")

    @property
    def label(self):
        return self._label

    def getItems(self, attribute=None):
        raise NotImplementedError("Override me")

    def drillUp(self, newAttribute):
        raise NotImplementedError("Override me")

    def dice(self, attribute, items, reorder=False):
        raise NotImplementedError("Override me")

    def diceRange(self, attribute, start, end):
        raise NotImplementedError("Override me")

    def getRootIndexFromRootItem(self, rootItem):
        return self.getItemsToIdx().get(rootItem, -1)

    def getGroupIndexFromRootIndex(self, groupAttr, rootIndex):
        raise NotImplementedError("Override me")

    def getItemsToIdx(self, attribute=None):
        attr = attribute or self._rootAttribute
        cached = self._itemsToIdx.get(attr)
        if cached is None:
            cached = {item: i for i, item in enumerate(self.getItems(attr))}
            self._itemsToIdx[attr] = cached
        return cached

    def getGroupIndexFromRootItem(self, groupAttr, rootItem):
        return self.getGroupIndexFromRootIndex(groupAttr, self.getRootIndexFromRootItem(rootItem))

    def getGroupItemFromRootIndex(self, groupAttr, rootIndex):
        return self.getItems(groupAttr)[self.getGroupIndexFromRootIndex(groupAttr, rootIndex)]

    def getGroupItemFromRootItem(self, groupAttr, rootItem):
        return self.getItems(groupAttr)[self.getGroupIndexFromRootItem(groupAttr, rootItem)]

    def _checkRootIndex(self, index):
        if index < 0 or index >= self.numItems:
            raise IndexError(f"rootIndex {index} out of bounds [0, {self.numItems}[")

    def _checkAttribute(self, attribute):
        if attribute not in self.attributes:
            raise KeyError(f"No attribute {attribute} was found on dimension {self.id}")
