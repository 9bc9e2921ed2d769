"""Process-wide registries that let the reference's call sequence stay unchanged
(testing_grounds.py:30-40) while the data stays off the text path."""

# sam_path -> BowtieRunner that aligned; PySamParser(sam_path).ranges reads its hit table
RESULTS = {}
# most recently constructed PAMFinder; BowtieRunner.align() fuses its PAM check into the search
ACTIVE_PAM = {"finder": None}
