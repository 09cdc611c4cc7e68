"""Logging set-up (reference: deepbelief/util.py:5-23): root logger to a file and to stdout."""
import logging
import sys


def init_logging(log_fname=None, level=logging.INFO):
    root = logging.getLogger()
    root.setLevel(level)
    fmt = logging.Formatter('%(asctime)s %(name)s %(levelname)s: %(message)s')
    if log_fname:
        fh = logging.FileHandler(log_fname)
        fh.setFormatter(fmt)
        root.addHandler(fh)
    sh = logging.StreamHandler(sys.stdout)
    sh.setFormatter(fmt)
    root.addHandler(sh)
    return root
