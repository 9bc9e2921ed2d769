"""Logging mixin with the reference's surface (Logger.py:10-94) and no babel/rich dependency.

Same levels (SUBPROC=25, HELP=15) and methods (info/debug/warn/error/subproc/help/json,
format_numbers).  Number formatting uses Python's locale-independent thousands grouping instead
of babel.format_decimal; rich is used for the console handler only when it is importable.
"""
import json
import logging


class Logger:
    SUBPROC = 25  # Between INFO (20) and WARNING (30)
    HELP = 15  # Between DEBUG (10) and INFO (20)
    _configured = False

    def __init__(self):
        if not Logger._configured:
            handlers = None
            try:
                from rich.console import Console
                from rich.logging import RichHandler
                handlers = [RichHandler(console=Console(stderr=True))]
            except Exception:  # rich is optional here
                handlers = [logging.StreamHandler()]
            logging.basicConfig(level=logging.INFO, format="%(message)s", datefmt="[%X]", handlers=handlers)
            logging.addLevelName(self.SUBPROC, "SUBPROC")
            logging.addLevelName(self.HELP, "HELP")
            Logger._configured = True
        self.logger = logging.getLogger("barcoder_b200")

    def format_numbers(self, message):
        if isinstance(message, str):
            lines = message.splitlines()
            for i, line in enumerate(lines):
                words = line.split()
                for j, word in enumerate(words):
                    try:
                        num = float(word)
                    except ValueError:
                        continue
                    words[j] = f"{int(num):,}" if num == int(num) and abs(num) < 1e18 else f"{num:,}"
                lines[i] = " ".join(words)
            message = "\n".join(lines)
        elif isinstance(message, int):
            message = f"{message:,}"
        return message

    def info(self, message):
        self.logger.info(self.format_numbers(message))

    def debug(self, message):
        self.logger.debug(self.format_numbers(message))

    def warn(self, message):
        self.logger.warning(self.format_numbers(message))

    def error(self, message):
        self.logger.error(self.format_numbers(message))

    def subproc(self, message, *args, **kwargs):
        message = self.format_numbers(message)
        if not message:
            message = "No errors reported"
        if self.logger.isEnabledFor(self.SUBPROC):
            self.logger._log(self.SUBPROC, message, args, **kwargs)

    def help(self, message, *args, **kwargs):
        message = self.format_numbers(message)
        if not message:
            message = "No help available"
        if self.logger.isEnabledFor(self.HELP):
            self.logger._log(self.HELP, message, args, **kwargs)

    def json(self, data):
        self.logger.info(json.dumps(data, indent=4))
