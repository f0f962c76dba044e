from .codes import Code
from .explanations import explain


class ModuleError(Exception):
    """Raised for the 4xxx codes; message is '<code>: <text>' (reference log/exceptions.py:25-27)."""

    def __init__(self, code: Code):
        self.code = code
        super().__init__(explain(code, show_code=True))
