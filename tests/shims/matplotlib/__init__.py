"""Stand-in for matplotlib (imported, never used on the path: R/utils/common.py:3).  Test infrastructure."""
