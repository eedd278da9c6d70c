"""Logger and progress-bar shim with the names the reference exposes (msm_we/_logging.py:7-42).
``rich`` is optional here: the hot-path functions only need ``add_task`` / ``update``."""
import logging

log = logging.getLogger("msm_we_b200")
if not log.handlers:
    log.addHandler(logging.NullHandler())


class _NullProgress:
    def add_task(self, description="", total=None, completed=0, **kw):
        return 0

    def update(self, task, advance=0, **kw):
        pass


class ProgressBar:
    """Context manager: uses the caller's progress object if one is given, else a no-op one."""

    def __init__(self, progress_bar=None):
        self.progress_bar = progress_bar if progress_bar is not None else _NullProgress()

    def __enter__(self):
        return self.progress_bar

    def __exit__(self, *exc):
        return False
