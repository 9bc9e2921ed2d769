"""barcoder_b200 - B200-native spacer->genome mismatch search behind barcoder's class API.

Host-side mirror of the reference interface for this ONE path (same module and class names):
BowtieRunner / BowtieError, PySamParser, PAMProcessor / PAMFinder / GuideFinder, CRISPRiLibrary,
GenBankParser / GenBankReader, BarCodeLibrary.  The compute path is the C-ABI CUDA library
(include/barcoder_b200.h, barcoder_b200/csrc); there is no CPU fallback.
"""
from .BarCodeLibrary import BarCodeLibrary, BarCodeLibraryError, BarCodeLibraryReader  # noqa: F401
from .BowtieRunner import BowtieError, BowtieRunner  # noqa: F401
from .CRISPRiLibrary import CRISPRiLibrary  # noqa: F401
from .GenBankParser import GenBankParser, GenBankReader  # noqa: F401
from .Logger import Logger  # noqa: F401
from .PAMProcessor import GuideFinder, PAMFinder, PAMProcessor  # noqa: F401
from .PySamParser import PySamParser  # noqa: F401

__version__ = "0.1.0"
