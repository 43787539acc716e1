"""Instrument tables the synthetiser's gain lookup reads.

Constants restated from reference ``utils/mapping_utils.py:56-106`` (GM-custom
pitch -> ADTOF class, class -> label) and the per-label gain table of
``modules/synthetiser.py:104-113``.  Stored as compact tables, not as the
reference's literal dicts.
"""
from __future__ import annotations

PITCH_MIN, PITCH_MAX = 35, 61  # reference synthetiser.py:252-253 (valid note range)

# GM-custom pitch 35..61 -> ADTOF class pitch (mapping_utils.py:56-84)
_ADTOF_CLASS_OF = (
    35, 35, 38, 38, 38, 38, 41, 42, 42, 42, 41, 48, 41, 48, 48, 42, 48, 52,
    61, 61, 61, 61, 61, 58, 61, 61, 61,
)
ADTOF_MAP = {p: c for p, c in zip(range(PITCH_MIN, PITCH_MAX + 1), _ADTOF_CLASS_OF)}

# class pitch -> label (mapping_utils.py:97-106)
ADTOF_LABEL = {35: "BD", 38: "SD", 41: "TT", 42: "HH", 48: "CY + RD",
               52: "Cowbell", 58: "Claves", 61: "Other"}

# class pitch -> member GM-custom pitches (mapping_utils.py:86-95); order matters
# because random.choice indexes into it.
ADTOF_INVERSE = {
    35: [35, 36], 38: [37, 38, 39, 40], 41: [41, 45, 47], 42: [42, 43, 44, 50],
    48: [46, 48, 49, 51], 52: [52], 58: [58], 61: [53, 54, 55, 56, 57, 59, 60],
}

# label -> mixing gain (synthetiser.py:104-113)
LABEL_GAIN = {"BD": 1.0, "SD": 1.0, "TT": 1.0, "HH": 0.7, "CY + RD": 0.7,
              "Cowbell": 0.7, "Claves": 0.7, "Other": 1.0}

# similarity threshold -> HDF5 sub-group names, best first (synthetiser.py:172-184)
SIMILARITY_GROUPS = ("gold", "100-90", "90-80", "80-70", "70-60", "60-50",
                     "50-40", "40-30", "30-20", "20-10", "10-0")


def instrument_gain(pitch: int, adtof_mapping: bool) -> float:
    """Gain of one instrument track in the final mix (synthetiser.py:152-153).

    Raises KeyError exactly where the reference's dict lookups would.
    """
    key = pitch if adtof_mapping else ADTOF_MAP[pitch]
    return LABEL_GAIN[ADTOF_LABEL[key]]
