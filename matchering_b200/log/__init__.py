from .codes import Code
from .exceptions import ModuleError
from .handlers import set_handlers, warning, info, debug, debug_line, debug_enabled
