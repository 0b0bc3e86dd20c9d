"""Per-file loop of `metadamage fit` (main.py:28-74): counts, then fits, for every valid file."""
import logging

from . import counts, fits, utils

logger = logging.getLogger(__name__)


def main(filenames, cfg):
    n_files = len(filenames)
    bad_files = 0
    for filename in filenames:
        if not utils.file_is_valid(filename):
            bad_files += 1
            continue
        cfg.add_filename(filename)
        df_counts = counts.load_counts(cfg)
        if not utils.is_df_counts_accepted(df_counts, cfg):
            continue
        fits.get_fits(df_counts, cfg)
        logger.debug("End of loop")
    if bad_files == n_files:
        raise Exception("All files were bad!")
