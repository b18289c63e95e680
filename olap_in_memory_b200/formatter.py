"""Flat <-> nested conversions of getData() results (host side, O(cells) but
presentation only).  Mirrors /root/reference/src/formatter/nested-array.js:1-37
and nested-object.js:1-47.  The reference's array helpers are only correct for
<= 2 dimensions (they re-read `values` on every loop turn, nested-array.js:6-8,
24-33); these handle any depth and agree with the reference for <= 2."""
from __future__ import annotations


def fromNestedArray(values, dimensions):
    flat = values
    for _ in range(len(dimensions) - 1):
        flat = [cell for row in flat for cell in row]
    return flat


def toNestedArray(values, _statusMap, dimensions):
    if len(dimensions) == 0:
        return values[0]
    nested = list(values)
    for dim in reversed(dimensions[1:]):
        chunk = dim.numItems
        nested = [nested[j * chunk:(j + 1) * chunk] for j in range(len(nested) // chunk)] if chunk else []
    return nested


def fromNestedObject(value, dimensions):
    level = [value]
    for dim in dimensions:
        items = dim.getItems()
        # a missing key reads as `undefined` in the reference, which unsets the cell
        level = [node.get(item) for node in level for item in items]
    return level


def toNestedObject(values, _statusMap, dimensions):
    def rec(depth, offset):
        if depth >= len(dimensions):
            return values[offset]
        items = dimensions[depth].getItems()
        return {item: rec(depth + 1, offset * len(items) + i) for i, item in enumerate(items)}

    return rec(0, 0)
