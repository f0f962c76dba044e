"""Process-wide log sinks: mg.log(default, warning_handler=..., info_handler=..., debug_handler=...,
show_codes=...) -- same call contract as the reference (matchering/log/handlers.py:24-83): the
handlers are plain callables taking one string; info/warning receive the Code's text."""
from .explanations import explain

_sinks = {"warning": None, "info": None, "debug": None}
_show_codes = False


def set_handlers(default_handler=None, warning_handler=None, info_handler=None, debug_handler=None,
                 show_codes=False):
    global _show_codes
    _sinks["warning"] = warning_handler or default_handler
    _sinks["info"] = info_handler or default_handler
    _sinks["debug"] = debug_handler or default_handler
    _show_codes = bool(show_codes)


def debug_enabled() -> bool:
    return _sinks["debug"] is not None


def warning(code):
    if _sinks["warning"]:
        _sinks["warning"](explain(code, _show_codes))


def info(code):
    if _sinks["info"]:
        _sinks["info"](explain(code, _show_codes))


def debug(message):
    """`message` may be a zero-argument callable so that expensive text (anything that needs a
    device read-back) is only built when somebody listens."""
    if _sinks["debug"]:
        _sinks["debug"](message() if callable(message) else message)


def debug_line():
    debug("-" * 40)
