"""Time-slot arithmetic for TimeDimension (host side, O(items), never O(cells)).

The reference delegates this to the third-party package ``timeslot-dag@2.2.0``
(/root/reference/package-lock.json:7284-7291), which is NOT vendored in the
reference tree.  This module restates its published behaviour for the call
sites the reference uses (src/dimension/time.js:1,7,19-20,55-61,71,92,118-131,
150-166,189-193):

  TimeSlot.fromValue, TimeSlot.fromDate, .periodicity, .value, .firstDate,
  .lastDate, .toParentPeriodicity, .next, TimeSlot.upperSlots

Parity is pinned only where the reference's own tests pin it
(test/dimension-time.js:11-88, test/cube-drilling.js:26-140,
test/cube-to-cube.js:168-183,553-605): day, month, quarter, semester, year,
all, week_mon, month_week_mon.  week_sat/week_sun/month_week_sat/_sun follow
the same published rule (epidemiological weeks: week 1 is the week holding at
least four days of the new year) and are parity-unpinned.

Dates are proleptic-Gregorian ordinals (``datetime.date``); the reference works
in UTC so there is no timezone component.
"""
from __future__ import annotations

import datetime as _dt
import re
from functools import lru_cache

_DAY = _dt.timedelta(days=1)

# python weekday(): Monday=0 .. Sunday=6
_WEEK_START = {"mon": 0, "sat": 5, "sun": 6}

UPPER_SLOTS = {
    "day": [
        "month_week_sat", "month_week_sun", "month_week_mon",
        "week_sat", "week_sun", "week_mon",
        "month", "quarter", "semester", "year", "all",
    ],
    "month_week_sat": ["week_sat", "month", "quarter", "semester", "year", "all"],
    "month_week_sun": ["week_sun", "month", "quarter", "semester", "year", "all"],
    "month_week_mon": ["week_mon", "month", "quarter", "semester", "year", "all"],
    "week_sat": ["month", "quarter", "semester", "year", "all"],
    "week_sun": ["month", "quarter", "semester", "year", "all"],
    "week_mon": ["month", "quarter", "semester", "year", "all"],
    "month": ["quarter", "semester", "year", "all"],
    "quarter": ["semester", "year", "all"],
    "semester": ["year", "all"],
    "year": ["all"],
    "all": [],
}

_RE = [
    ("day", re.compile(r"^(\d{4})-(\d{2})-(\d{2})$")),
    ("month_week", re.compile(r"^(\d{4})-(\d{2})-W(\d)-(sat|sun|mon)$")),
    ("week", re.compile(r"^(\d{4})-W(\d{2})-(sat|sun|mon)$")),
    ("month", re.compile(r"^(\d{4})-(\d{2})$")),
    ("quarter", re.compile(r"^(\d{4})-Q(\d)$")),
    ("semester", re.compile(r"^(\d{4})-S(\d)$")),
    ("year", re.compile(r"^(\d{4})$")),
]


def _month_end(year: int, month: int) -> _dt.date:
    if month == 12:
        return _dt.date(year, 12, 31)
    return _dt.date(year, month + 1, 1) - _DAY


def _week_epoch(year: int, start: str) -> _dt.date:
    """First day of week 1 of `year`: the week (starting on `start`) that
    contains January 4th, i.e. holds at least four days of the new year."""
    jan4 = _dt.date(year, 1, 4)
    back = (jan4.weekday() - _WEEK_START[start]) % 7
    return jan4 - back * _DAY


def _first_month_week_length(year: int, month: int, start: str) -> int:
    """Number of days of the (possibly partial) first week of the month."""
    first = _dt.date(year, month, 1)
    # days until the next week start (a month beginning on the week start has a full first week)
    return 7 - ((first.weekday() - _WEEK_START[start]) % 7)


_MONTHS_EN = ["January", "February", "March", "April", "May", "June", "July", "August", "September", "October",
              "November", "December"]


class TimeSlot:
    """Immutable time slot; compare/sort by ``.value`` (a string), like the reference."""

    upperSlots = UPPER_SLOTS

    __slots__ = ("value", "periodicity", "_first", "_last")

    def __init__(self, value: str):
        self.value = value
        self._first = None
        self._last = None
        if value == "all":
            self.periodicity = "all"
            return
        for kind, rx in _RE:
            m = rx.match(value)
            if m:
                if kind == "month_week":
                    self.periodicity = "month_week_" + m.group(4)
                elif kind == "week":
                    self.periodicity = "week_" + m.group(3)
                else:
                    self.periodicity = kind
                return
        raise ValueError(f"Invalid time slot: {value}")

    # ---- construction -------------------------------------------------
    @staticmethod
    @lru_cache(maxsize=65536)
    def fromValue(value: str) -> "TimeSlot":
        return TimeSlot(value)

    @staticmethod
    def fromDate(date: _dt.date, periodicity: str) -> "TimeSlot":
        y, mth, d = date.year, date.month, date.day
        if periodicity == "day":
            return TimeSlot.fromValue(f"{y:04d}-{mth:02d}-{d:02d}")
        if periodicity.startswith("month_week_"):
            start = periodicity[-3:]
            fwl = _first_month_week_length(y, mth, start)
            week = 1 if d <= fwl else (d - 1 - fwl) // 7 + 2
            return TimeSlot.fromValue(f"{y:04d}-{mth:02d}-W{week}-{start}")
        if periodicity.startswith("week_"):
            start = periodicity[-3:]
            year = y + 1
            epoch = _week_epoch(year, start)
            while date < epoch:
                year -= 1
                epoch = _week_epoch(year, start)
            week = (date - epoch).days // 7 + 1
            return TimeSlot.fromValue(f"{year:04d}-W{week:02d}-{start}")
        if periodicity == "month":
            return TimeSlot.fromValue(f"{y:04d}-{mth:02d}")
        if periodicity == "quarter":
            return TimeSlot.fromValue(f"{y:04d}-Q{1 + (mth - 1) // 3}")
        if periodicity == "semester":
            return TimeSlot.fromValue(f"{y:04d}-S{1 + (mth - 1) // 6}")
        if periodicity == "year":
            return TimeSlot.fromValue(f"{y:04d}")
        if periodicity == "all":
            return TimeSlot.fromValue("all")
        raise ValueError(f"Invalid periodicity: {periodicity}")

    # ---- bounds --------------------------------------------------------
    @property
    def firstDate(self) -> _dt.date:
        if self._first is None:
            self._compute_bounds()
        return self._first

    @property
    def lastDate(self) -> _dt.date:
        if self._last is None:
            self._compute_bounds()
        return self._last

    def _compute_bounds(self) -> None:
        p, v = self.periodicity, self.value
        if p == "day":
            self._first = self._last = _dt.date(int(v[0:4]), int(v[5:7]), int(v[8:10]))
        elif p.startswith("month_week_"):
            y, mth, week, start = int(v[0:4]), int(v[5:7]), int(v[9]), v[-3:]
            fwl = _first_month_week_length(y, mth, start)
            first_day = 1 if week == 1 else 1 + fwl + (week - 2) * 7
            last_day = fwl if week == 1 else first_day + 6
            end = _month_end(y, mth)
            self._first = _dt.date(y, mth, first_day)
            self._last = min(_dt.date(y, mth, min(last_day, end.day)), end)
        elif p.startswith("week_"):
            y, week, start = int(v[0:4]), int(v[6:8]), v[-3:]
            self._first = _week_epoch(y, start) + (week - 1) * 7 * _DAY
            self._last = self._first + 6 * _DAY
        elif p == "month":
            y, mth = int(v[0:4]), int(v[5:7])
            self._first, self._last = _dt.date(y, mth, 1), _month_end(y, mth)
        elif p == "quarter":
            y, q = int(v[0:4]), int(v[6])
            self._first, self._last = _dt.date(y, 3 * q - 2, 1), _month_end(y, 3 * q)
        elif p == "semester":
            y, s = int(v[0:4]), int(v[6])
            self._first, self._last = _dt.date(y, 6 * s - 5, 1), _month_end(y, 6 * s)
        elif p == "year":
            y = int(v)
            self._first, self._last = _dt.date(y, 1, 1), _dt.date(y, 12, 31)
        else:
            raise ValueError("'all' has no bounds")

    # ---- navigation ----------------------------------------------------
    def toParentPeriodicity(self, periodicity: str) -> "TimeSlot":
        if periodicity == self.periodicity:
            return self
        if periodicity not in UPPER_SLOTS[self.periodicity]:
            raise ValueError(f"Cannot convert {self.periodicity} to {periodicity}")
        if periodicity == "all":
            return TimeSlot.fromValue("all")
        date = self.firstDate
        # a week belongs to the month/quarter/year of its middle day
        if self.periodicity in ("week_sat", "week_sun", "week_mon"):
            date = date + 3 * _DAY
        return TimeSlot.fromDate(date, periodicity)

    def next(self) -> "TimeSlot":
        return TimeSlot.fromDate(self.lastDate + _DAY, self.periodicity)

    def previous(self) -> "TimeSlot":
        return TimeSlot.fromDate(self.firstDate - _DAY, self.periodicity)

    def humanizeValue(self, language: str = "en") -> str:
        """Labels of timeslot-dag's `humanizeValue`.  Only the forms the reference's tests pin
        are restated (test/dimension-time.js:186-222: English months, French quarters); every
        other periodicity / language falls back to the slot value itself (unpinned)."""
        if self.periodicity == "month" and language == "en":
            year, month = self.value.split("-")
            return f"{_MONTHS_EN[int(month) - 1]} {year}"
        if self.periodicity == "quarter" and language == "fr":
            year, quarter = self.value.split("-Q")
            return f"{'1er' if quarter == '1' else quarter + 'ème'} trim. {year}"
        return self.value

    def __repr__(self) -> str:
        return f"TimeSlot({self.value!r})"
