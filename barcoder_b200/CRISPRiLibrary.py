"""CRISPRiLibrary with the reference's surface (CRISPRiLibrary.py:4-120): takes the joined
hit x feature frame and a PAMFinder, exposes targets_df, source_unique_targets, mapped_targets,
unique_targets and unambiguous_targets.

Differences in mechanism, not in result:
  * `_annotate_targets` copies the PAM / Targeting columns the CUDA search already produced when
    they were computed for this finder's PAM; only otherwise does it fall back to the finder's
    per-row string functions (CRISPRiLibrary.py:16-21);
  * Offset / Overlap are column arithmetic instead of row-wise apply (CRISPRiLibrary.py:61-83).
"""
import numpy as np
import pandas as pd


class CRISPRiLibrary:
    def __init__(self, pyranges_df, pam_finder):
        self.targets_df = pyranges_df
        self.pam_finder = pam_finder
        self._annotate_targets()
        self.source_unique_targets = self._get_source_unique_targets()
        self.mapped_targets = self._get_mapped_targets()
        self.unique_targets = self._get_unique_targets()
        self.unambiguous_targets = self._get_unambiguous_targets()

    def _annotate_targets(self):
        df, finder = self.targets_df, self.pam_finder
        fused = {"PAM", "Targeting"} <= set(df.columns) and \
            df.attrs.get("pam_key") == (str(getattr(finder, "raw_pam", "")).upper(), "class-api")
        if fused:
            return
        df["PAM"] = [finder.get_pam_seq(row) for row in df.itertuples(index=False)] if len(df) else []
        df["Targeting"] = [finder.pam_matches(s) for s in df["PAM"]] if len(df) else []

    def _targeting_mapped(self):
        df = self.targets_df
        return df["Targeting"].astype(bool) & df["Mapped"].astype(bool)

    def _get_source_unique_targets(self):
        """Barcodes with exactly... the FIRST `source`-feature row of every targeting, mapped
        barcode (rows whose Barcode was already seen are dropped; CRISPRiLibrary.py:37-45)."""
        df = self.targets_df
        sel = df[(df["Type"] == "source") & self._targeting_mapped()]
        return sel[~sel.duplicated(subset=["Barcode"])].reset_index(drop=True)

    def _get_mapped_targets(self):
        df = self.targets_df
        sel = df[(df["Type"] != "source") & self._targeting_mapped()].copy()
        start, end = sel["Start"].to_numpy(dtype=np.int64), sel["End"].to_numpy(dtype=np.int64)
        fs, fe = sel["Start_b"].to_numpy(dtype=np.int64), sel["End_b"].to_numpy(dtype=np.int64)
        strand_b = sel["Strand_b"].to_numpy()
        offset = np.where(strand_b == "+", start - fs, fe - end).astype(object)
        offset[(strand_b != "+") & (strand_b != "-")] = None
        sel["Offset"] = offset if len(sel) else []
        sel["Overlap"] = np.maximum(np.minimum(end, fe) - np.maximum(start, fs), 0)
        return sel.reset_index(drop=True)

    def _get_unique_targets(self):
        mapped = self._get_mapped_targets()
        keep = mapped["Barcode"].isin(self.source_unique_targets.Barcode)
        return mapped[keep].sort_values(["Chromosome", "Start", "End"]).reset_index(drop=True)

    def _get_unambiguous_targets(self):
        ut = self.unique_targets
        return ut[~ut.duplicated(subset=["Barcode"]).reset_index(drop=True)]
