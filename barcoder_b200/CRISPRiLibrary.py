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
        # barcodes as integer codes: duplicated / isin on 10^6 strings cost seconds, on codes milliseconds
        self._codes = pd.factorize(self.targets_df["Barcode"])[0] if len(self.targets_df) else np.zeros(0, dtype=np.int64)
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

    def _first_of_each_barcode(self, index):
        """Boolean array over `index` (row labels of targets_df): True where the row is the first one of its barcode."""
        codes = self._codes[np.asarray(index)]
        first = np.zeros(len(codes), dtype=bool)
        first[np.unique(codes, return_index=True)[1]] = True
        return first

    def _get_source_unique_targets(self):
        """Barcodes with exactly... the FIRST `source`-feature row of every targeting, mapped
        barcode (rows whose Barcode was already seen are dropped; CRISPRiLibrary.py:37-45)."""
        df = self.targets_df
        rows = np.nonzero((df["Type"] == "source").to_numpy() & self._targeting_mapped().to_numpy())[0]
        return df.iloc[rows[self._first_of_each_barcode(rows)]].reset_index(drop=True)  # one take of the string columns

    def _get_mapped_targets(self):
        df = self.targets_df
        sel = df[(df["Type"] != "source").to_numpy() & self._targeting_mapped().to_numpy()].copy()
        self._mapped_rows = np.asarray(sel.index)
        start, end = sel["Start"].to_numpy(dtype=np.int64), sel["End"].to_numpy(dtype=np.int64)
        fs, fe = sel["Start_b"].to_numpy(dtype=np.int64), sel["End_b"].to_numpy(dtype=np.int64)
        strand_b = sel["Strand_b"].to_numpy()
        offset = np.where(strand_b == "+", start - fs, fe - end).astype(object)
        offset[(strand_b != "+") & (strand_b != "-")] = None
        sel["Offset"] = offset if len(sel) else []
        sel["Overlap"] = np.maximum(np.minimum(end, fe) - np.maximum(start, fs), 0)
        return sel.reset_index(drop=True)

    def _get_unique_targets(self):
        mapped = self.mapped_targets          # built once in __init__ (CRISPRiLibrary.py:86 recomputes it; same rows)
        df = self.targets_df
        src = (df["Type"] == "source").to_numpy() & self._targeting_mapped().to_numpy()
        in_source = np.zeros(int(self._codes.max()) + 1 if len(self._codes) else 0, dtype=bool)
        in_source[self._codes[src]] = True    # Barcode.isin(source_unique_targets.Barcode)
        keep = np.nonzero(in_source[self._codes[self._mapped_rows]])[0] if len(mapped) else np.zeros(0, dtype=np.int64)
        # sort keys from the columns, then ONE take of the frame (every take copies ~8 string columns)
        # (End is an object column when the hit frame carries unmapped reads; as int64 the sort is 20x faster)
        order = np.lexsort((mapped["End"].to_numpy(dtype=np.int64)[keep], mapped["Start"].to_numpy(dtype=np.int64)[keep],
                            pd.factorize(mapped["Chromosome"], sort=True)[0][keep])) if len(keep) else np.zeros(0, dtype=np.int64)
        rows = keep[order]
        self._unique_codes = self._codes[self._mapped_rows][rows]
        return mapped.iloc[rows].reset_index(drop=True)

    def _get_unambiguous_targets(self):
        ut = self.unique_targets
        if not len(ut):
            return ut
        first = np.zeros(len(ut), dtype=bool)
        first[np.unique(self._unique_codes, return_index=True)[1]] = True
        return ut[first]
