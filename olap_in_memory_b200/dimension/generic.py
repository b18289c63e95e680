"""Dimension whose items are arbitrary strings, with optional parent attributes
(city -> country -> continent ...).  Host-side mirror of
/root/reference/src/dimension/generic.js:4-345.

What matters for the device path is the root-item -> group-item index map of
every attribute (generic.js:83-113: groups are numbered by first appearance
while walking root items in order; the virtual attribute ``all`` maps every
item to 0, generic.js:30).  Those maps are kept as int32 numpy arrays and are
handed to the C ABI unchanged."""
from __future__ import annotations

import numpy as np

from .abstract import AbstractDimension


def _get_or_call(objfun, param):
    if objfun is None:  # JS `!objfun`: an empty object {} is truthy there (labels become undefined, not the item)
        return param
    if callable(objfun):
        return objfun(param)
    return objfun.get(param)


class GenericDimension(AbstractDimension):
    def __init__(self, id, rootAttribute, items, label=None, itemToLabelMap=None):
        super().__init__(id, rootAttribute, label)
        items = list(items)
        self._items = {"all": ["all"], rootAttribute: items}
        self._rootIdxToGroupIdx = {
            "all": np.zeros(len(items), dtype=np.int32),
            rootAttribute: np.arange(len(items), dtype=np.int32),
        }
        self._itemToLabel = {
            "all": {"all": "All"},
            rootAttribute: {item: _get_or_call(itemToLabelMap, item) for item in items},
        }

    @property
    def attributes(self):
        return list(self._rootIdxToGroupIdx.keys())

    @staticmethod
    def deserialize(buffer):  # generic.js:47-60
        from ..serialization import fromBuffer

        data = fromBuffer(buffer)
        dim = GenericDimension(data["id"], data["rootAttribute"], data["attributeItems"][data["rootAttribute"]],
                               data["label"])
        dim._items.update(data["attributeItems"])
        dim._itemToLabel.update(data["attributeLabels"])
        dim._rootIdxToGroupIdx.update({k: np.asarray(v, dtype=np.int32) for k, v in data["attributeMappings"].items()})
        return dim

    def serialize(self):  # generic.js:62-72
        from ..serialization import toBuffer

        return toBuffer({
            "id": self.id,
            "label": self.label,
            "rootAttribute": self._rootAttribute,
            "rootItems": self._items[self._rootAttribute],
            "attributeItems": self._items,
            "attributeLabels": self._itemToLabel,
            "attributeMappings": {k: np.asarray(v).astype(np.uint32) for k, v in self._rootIdxToGroupIdx.items()},
        })

    def addAttribute(self, baseAttr, newAttr, baseToNew, newToNewLabel=None):
        """Derive attribute `newAttr` from `baseAttr` through `baseToNew`
        (dict or callable).  Nothing is modified if the mapping raises."""
        new_index = {}
        items = []
        labels = {}
        base_map = self._rootIdxToGroupIdx[baseAttr]
        base_items = self._items[baseAttr]
        mapping = np.empty(self.numItems, dtype=np.int32)
        for root in range(self.numItems):
            new_item = _get_or_call(baseToNew, base_items[base_map[root]])
            if not isinstance(new_item, str):
                raise TypeError("Mapping result must be a string.")
            slot = new_index.get(new_item)
            if slot is None:
                slot = new_index[new_item] = len(items)
                items.append(new_item)
                labels[new_item] = _get_or_call(newToNewLabel, new_item)
            mapping[root] = slot
        self._items[newAttr] = items
        self._rootIdxToGroupIdx[newAttr] = mapping
        self._itemToLabel[newAttr] = labels

    def getItems(self, attribute=None):
        return self._items[attribute or self._rootAttribute]

    def getEntries(self, attribute=None):
        attr = attribute or self._rootAttribute
        return [[item, self._itemToLabel[attr][item]] for item in self._items[attr]]

    def renameItem(self, oldItem, newItem, newLabel=None):
        if newItem in self._items[self._rootAttribute]:
            raise ValueError(f"Item {newItem} already exists")
        for items in self._items.values():
            if oldItem in items:
                items[items.index(oldItem)] = newItem
        for idx in self._itemsToIdx.values():
            if oldItem in idx:
                idx[newItem] = idx.pop(oldItem)
        for labels in self._itemToLabel.values():
            if labels.get(oldItem):
                labels[newItem] = newLabel or newItem
                del labels[oldItem]

    def drillUp(self, targetAttr):
        if targetAttr == self._rootAttribute:
            return self
        new_dim = GenericDimension(
            self.id, targetAttr, self.getItems(targetAttr), self.label, self._itemToLabel[targetAttr]
        )
        new_items = self._items[targetAttr]
        new_map = self._rootIdxToGroupIdx[targetAttr]
        n_root = len(self._items[self._rootAttribute])
        for child_attr in self.attributes:
            if child_attr == targetAttr:
                continue
            child_items = self._items[child_attr]
            child_map = self._rootIdxToGroupIdx[child_attr]
            parent_of = {}
            clean_cut = True
            for i in range(n_root):
                child_item = child_items[child_map[i]]
                new_item = new_items[new_map[i]]
                # the attribute survives only if every new item has a single parent
                if parent_of.get(new_item) and parent_of[new_item] != child_item:
                    clean_cut = False
                    break
                parent_of[new_item] = child_item
            if clean_cut:
                new_dim.addAttribute(targetAttr, child_attr, parent_of, self._itemToLabel[child_attr])
        return new_dim

    def dice(self, attribute, items, reorder=False):
        old_items = self._items[self._rootAttribute]
        if self._rootAttribute == attribute:
            if reorder:
                new_items = [i for i in items if i in old_items]
            else:
                wanted = set(items)
                new_items = [i for i in old_items if i in wanted]
        else:
            if reorder:
                raise ValueError("Reordering is not allowed when using groups")
            wanted = set(items)
            new_items = [i for i in old_items if self.getGroupItemFromRootItem(attribute, i) in wanted]

        if new_items == old_items:
            return self

        dimension = GenericDimension(
            self.id, self._rootAttribute, new_items, self.label, self._itemToLabel[self._rootAttribute]
        )
        for attr in self.attributes:
            if attr != self._rootAttribute:
                dimension.addAttribute(
                    self._rootAttribute,
                    attr,
                    lambda item, attr=attr: self.getGroupItemFromRootItem(attr, item),
                    self._itemToLabel[attr],
                )
        return dimension

    def getGroupIndexFromRootIndexMap(self, groupAttr):
        self._checkAttribute(groupAttr)
        return self._rootIdxToGroupIdx[groupAttr]

    def getGroupIndexFromRootIndex(self, groupAttr, rootIdx):
        self._checkAttribute(groupAttr)
        self._checkRootIndex(rootIdx)
        return int(self._rootIdxToGroupIdx[groupAttr][rootIdx])

    def union(self, otherDimension):
        if self.id != otherDimension.id:
            raise ValueError("not the same dimension")
        me, other = self, otherDimension
        if otherDimension._rootAttribute in self.attributes:
            me = me.drillUp(otherDimension._rootAttribute)
        elif self._rootAttribute in otherDimension.attributes:
            other = other.drillUp(self._rootAttribute)
        else:
            raise ValueError("The dimensions are not compatible")

        def any_item_to_group(group_attr, root_item):
            try:
                return me.getGroupItemFromRootItem(group_attr, root_item)
            except Exception:
                return other.getGroupItemFromRootItem(group_attr, root_item)

        def any_item_to_label(attr, item):
            for dim in (me, other):
                found = dim._itemToLabel.get(attr, {}).get(item)
                if found:
                    return found
            return item

        mine = me.getItems()
        merged = sorted(list(mine) + [i for i in otherDimension.getItems() if i not in mine])
        dimension = GenericDimension(
            me.id, me._rootAttribute, merged, me.label, lambda item: any_item_to_label(me._rootAttribute, item)
        )
        group_attrs = [a for a in me.attributes if a != me._rootAttribute]
        group_attrs += [a for a in other.attributes if a != other._rootAttribute and a not in group_attrs]
        for group_attr in group_attrs:
            try:
                dimension.addAttribute(
                    me._rootAttribute,
                    group_attr,
                    lambda root_item, g=group_attr: any_item_to_group(g, root_item),
                    lambda group_item, g=group_attr: any_item_to_label(g, group_item),
                )
            except Exception:
                pass
        return dimension

    def intersect(self, otherDimension):
        if self.id != otherDimension.id:
            raise ValueError("not the same dimension")
        if otherDimension._rootAttribute in self.attributes:
            root = otherDimension._rootAttribute
        elif self._rootAttribute in otherDimension.attributes:
            root = self._rootAttribute
        else:
            raise ValueError("The dimensions are not compatible")
        other_items = set(otherDimension.getItems(root))
        common = [i for i in self.getItems(root) if i in other_items]
        return self.drillUp(root).dice(root, common)
