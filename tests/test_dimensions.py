"""Index-map producers: the reference's own dimension tests, transcribed
(test/dimension-generic.js, test/dimension-time.js).  Serialization and the humanised time
labels (timeslot-dag's i18n) are outside the path and are not transcribed."""
import numpy as np
import pytest

from olap_in_memory_b200 import GenericDimension, TimeDimension


@pytest.fixture
def generic():  # dimension-generic.js:7-45
    d = GenericDimension("location", "city", ["paris", "toulouse", "madrid", "beirut"], "Location",
                         lambda item: f"city of {item}")
    d.addAttribute("city", "cityNumLetters", lambda city: str(len(city)), {"5": "five", "6": "six", "8": "eigth"})
    d.addAttribute("city", "country", {"madrid": "spain", "beirut": "lebanon", "paris": "france", "toulouse": "france"},
                   lambda item: f"country of {item}")
    d.addAttribute("country", "continent", lambda item: "asia" if item == "lebanon" else "europe",
                   {"asia": "The huge continent", "europe": "The old continent"})
    return d


def test_generic_sizes_attributes_items(generic):  # :47-78
    assert generic.numItems == 4
    assert generic.rootAttribute == "city"
    assert sorted(generic.attributes) == sorted(["city", "cityNumLetters", "country", "continent", "all"])
    assert generic.getItems() == ["paris", "toulouse", "madrid", "beirut"]
    assert generic.getItems("city") == ["paris", "toulouse", "madrid", "beirut"]
    assert generic.getItems("cityNumLetters") == ["5", "8", "6"]


def test_generic_group_items_and_indexes(generic):  # :80-99
    assert generic.getGroupItemFromRootItem("city", "paris") == "paris"
    assert generic.getGroupItemFromRootItem("cityNumLetters", "madrid") == "6"
    assert generic.getGroupItemFromRootItem("country", "madrid") == "spain"
    assert generic.getGroupItemFromRootItem("continent", "madrid") == "europe"
    assert [generic.getGroupIndexFromRootIndex("country", i) for i in range(4)] == [0, 0, 1, 2]
    # the arrays that cross the C ABI: int32, groups numbered by first appearance (generic.js:83-113)
    m = generic.getGroupIndexFromRootIndexMap("country")
    assert m.dtype == np.int32 and m.tolist() == [0, 0, 1, 2]
    assert generic.getGroupIndexFromRootIndexMap("all").tolist() == [0, 0, 0, 0]
    assert generic.getGroupIndexFromRootIndexMap("cityNumLetters").tolist() == [0, 1, 2, 2]


def test_generic_drill_up(generic):  # :101-108
    child = generic.drillUp("country")
    assert sorted(child.attributes) == sorted(["country", "continent", "all"])
    assert child.getItems() == ["france", "spain", "lebanon"]
    assert child.getGroupIndexFromRootIndexMap("continent").tolist() == [0, 0, 1]
    assert sorted(generic.drillUp("cityNumLetters").attributes) == sorted(["cityNumLetters", "all"])


def test_generic_intersect_and_union(generic):  # :110-191
    other = GenericDimension("location", "city", ["toulouse", "madrid", "amman", "paris"])
    inter = generic.intersect(other)
    assert inter.rootAttribute == "city" and inter.getItems() == ["paris", "toulouse", "madrid"]
    inter = generic.intersect(GenericDimension("location", "country", ["france", "spain", "jordan"]))
    assert inter.rootAttribute == "country" and inter.getItems() == ["france", "spain"]
    empty = generic.intersect(GenericDimension("location", "city", ["lyon", "barcelona", "narbonne"]))
    assert empty.numItems == 0 and empty.getItems() == []
    with pytest.raises(Exception):
        generic.intersect(GenericDimension("location", "postalcode", ["75018", "75019"]))
    other = GenericDimension("location", "city", ["lyon"], "Location", lambda item: f"great city of {item}")
    other.addAttribute("city", "country", lambda _c: "france", lambda item: f"country of {item}")
    union = generic.union(other)
    assert union.attributes == ["all", "city", "country"]
    assert union.getGroupItemFromRootItem("country", "lyon") == "france"
    assert union.getGroupItemFromRootItem("country", "paris") == "france"
    assert union.getEntries() == [["beirut", "city of beirut"], ["lyon", "great city of lyon"],
                                  ["madrid", "city of madrid"], ["paris", "city of paris"],
                                  ["toulouse", "city of toulouse"]]


def test_generic_labels(generic):  # :199-237
    assert generic.getEntries() == [["paris", "city of paris"], ["toulouse", "city of toulouse"],
                                    ["madrid", "city of madrid"], ["beirut", "city of beirut"]]
    assert generic.getEntries("cityNumLetters") == [["5", "five"], ["8", "eigth"], ["6", "six"]]
    assert generic.drillUp("cityNumLetters").getEntries() == [["5", "five"], ["8", "eigth"], ["6", "six"]]
    diced = generic.dice("cityNumLetters", ["6", "5"])
    assert diced.getEntries() == [["paris", "city of paris"], ["madrid", "city of madrid"], ["beirut", "city of beirut"]]
    assert diced.getEntries("cityNumLetters") == [["5", "five"], ["6", "six"]]


@pytest.fixture
def months():  # dimension-time.js:7-9
    return TimeDimension("time", "month", "2009-12", "2010-02")


def test_time_sizes_attributes_items(months):  # :11-36
    assert months.numItems == 3 and months.rootAttribute == "month"
    assert sorted(months.attributes) == sorted(["month", "quarter", "semester", "year", "all"])
    assert months.getItems() == ["2009-12", "2010-01", "2010-02"]
    assert months.getItems("month") == ["2009-12", "2010-01", "2010-02"]
    assert months.getItems("year") == ["2009", "2010"]


def test_time_group_items_and_indexes(months):  # :38-50
    assert months.getGroupItemFromRootItem("month", "2010-01") == "2010-01"
    assert months.getGroupItemFromRootItem("year", "2010-01") == "2010"
    assert [months.getGroupIndexFromRootIndex("month", i) for i in (0, 1)] == [0, 1]
    assert [months.getGroupIndexFromRootIndex("year", i) for i in (0, 1)] == [0, 1]
    m = months.getGroupIndexFromRootIndexMap("year")
    assert m.dtype == np.int32 and m.tolist() == [0, 1, 1]
    assert months.getGroupIndexFromRootIndexMap("quarter").tolist() == [0, 1, 1]
    assert months.getGroupIndexFromRootIndexMap("all").tolist() == [0, 0, 0]


def test_time_drill_up_and_down(months):  # :52-88
    child = months.drillUp("quarter")
    assert sorted(child.attributes) == sorted(["quarter", "semester", "year", "all"])
    assert child.getItems() == ["2009-Q4", "2010-Q1"]
    weeks = months.drillDown("week_mon")
    assert sorted(weeks.attributes) == sorted(["week_mon", "month", "quarter", "semester", "year", "all"])
    assert weeks.getItems() == ["2009-W49-mon", "2009-W50-mon", "2009-W51-mon", "2009-W52-mon", "2009-W53-mon",
                                "2010-W01-mon", "2010-W02-mon", "2010-W03-mon", "2010-W04-mon", "2010-W05-mon",
                                "2010-W06-mon", "2010-W07-mon", "2010-W08-mon"]


def test_time_intersect_union(months):  # :90-146
    inter = months.intersect(TimeDimension("time", "month", "2010-01", "2010-02"))
    assert inter.rootAttribute == "month" and inter.getItems() == ["2010-01", "2010-02"]
    inter = months.intersect(TimeDimension("time", "quarter", "2010-Q1", "2010-Q2"))
    assert inter.rootAttribute == "quarter" and inter.getItems() == ["2010-Q1"]
    empty = months.intersect(TimeDimension("time", "quarter", "2010-Q3", "2010-Q4"))
    assert empty.numItems == 0 and empty.getItems() == []
    union = months.union(TimeDimension("time", "quarter", "2010-Q3", "2010-Q4"))
    assert union.rootAttribute == "quarter"
    assert union.getItems() == ["2009-Q4", "2010-Q1", "2010-Q2", "2010-Q3", "2010-Q4"]


def test_time_dice_range(months):  # :157-181
    assert months.diceRange("month", "2010-01", "2010-01").getItems() == ["2010-01"]
    assert months.diceRange("month", "2000-01", "2020-01").getItems() == ["2009-12", "2010-01", "2010-02"]
    assert months.diceRange("month", "2010-01", "2020-01").getItems() == ["2010-01", "2010-02"]
    assert months.diceRange("month", "2010-01", None).getItems() == ["2010-01", "2010-02"]
    assert months.diceRange("month", None, "2010-01").getItems() == ["2009-12", "2010-01"]
    assert months.dice("quarter", ["2010-Q1"]).getItems() == ["2010-01", "2010-02"]


def test_config_maps_day_to_month():
    """The map every BASELINE config uses: 3652 days -> 120 months, monotone, 28-31 per group."""
    day = TimeDimension("time", "day", "2010-01-01", "2019-12-31")
    m = day.getGroupIndexFromRootIndexMap("month")
    assert day.numItems == 3652 and m.dtype == np.int32
    assert bool(np.all(np.diff(m) >= 0)) and m[0] == 0 and m[-1] == 119
    counts = np.bincount(m)
    assert counts.tolist()[:3] == [31, 28, 31] and counts[25] == 29  # 2012-02 is a leap February
    assert set(counts.tolist()) == {28, 29, 30, 31}
    assert day.getGroupIndexFromRootIndexMap("year").max() == 9


def test_generic_serialized(generic):  # dimension-generic.js:191-195
    new = GenericDimension.deserialize(generic.serialize())
    assert new.getItems() == generic.getItems()
    for attribute in generic.attributes:
        assert new.getItems(attribute) == generic.getItems(attribute)
        assert new.getGroupIndexFromRootIndexMap(attribute).tolist() == generic.getGroupIndexFromRootIndexMap(attribute).tolist()


def test_time_serialized(months):  # dimension-time.js:148-155
    new = TimeDimension.deserialize(months.serialize())
    assert months.getItems() == new.getItems()
    assert months.getItems("quarter") == new.getItems("quarter")


def test_time_labels(months):  # dimension-time.js:186-222
    assert months.getEntries() == [["2009-12", "December 2009"], ["2010-01", "January 2010"], ["2010-02", "February 2010"]]
    assert months.getEntries("quarter", "fr") == [["2009-Q4", "4ème trim. 2009"], ["2010-Q1", "1er trim. 2010"]]
    assert months.drillUp("quarter").getEntries(None, "fr") == [["2009-Q4", "4ème trim. 2009"], ["2010-Q1", "1er trim. 2010"]]
    diced = months.dice("quarter", ["2010-Q1"])
    assert diced.getEntries() == [["2010-01", "January 2010"], ["2010-02", "February 2010"]]
    assert diced.getEntries("quarter", "fr") == [["2010-Q1", "1er trim. 2010"]]
