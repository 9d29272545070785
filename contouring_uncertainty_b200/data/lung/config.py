"""Point counts and labels of the lung / heart contours (contour_uncertainty/data/lung/config.py:9-42)."""
import enum


class Label(enum.IntEnum):
    BG = 0
    LUNG = 1
    HEART = 2


RLUNG = 44
LLUNG = 50
HEART = 26
