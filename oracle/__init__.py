"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/concrete.py)."""
