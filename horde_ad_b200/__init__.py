"""Import shim: the product package lives in the directory `horde-ad_b200/`
(the repository's contractual name, which is not a valid Python identifier).
This package extends its search path to that directory so that
`import horde_ad_b200.device` etc. resolve to horde-ad_b200/*.py.
"""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "horde-ad_b200"))
