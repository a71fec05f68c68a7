"""setup_logger with the signature of the reference's utils/logger_utils.py:5 (name, log_dir, console_level, file_level,
console_format_str, file_format_str): a named logger writing to the console and to <log_dir>/<name>.log."""
import logging
import os

_DEFAULT_FORMAT = "%(asctime)s - %(name)s - %(levelname)s - %(message)s"


def setup_logger(name, log_dir, console_level=logging.INFO, file_level=logging.DEBUG, console_format_str=None,
                 file_format_str=None):
    logger = logging.getLogger(name)
    logger.setLevel(logging.DEBUG)
    if logger.handlers:  # configured before (the reference guards the same way: handlers are added once per name)
        return logger
    if console_level is not None:
        console = logging.StreamHandler()
        console.setLevel(console_level)
        console.setFormatter(logging.Formatter(console_format_str or _DEFAULT_FORMAT))
        logger.addHandler(console)
    if file_level is not None and log_dir:
        os.makedirs(log_dir, exist_ok=True)
        to_file = logging.FileHandler(os.path.join(log_dir, f"{name}.log"))
        to_file.setLevel(file_level)
        to_file.setFormatter(logging.Formatter(file_format_str or _DEFAULT_FORMAT))
        logger.addHandler(to_file)
    logger.propagate = False
    return logger
