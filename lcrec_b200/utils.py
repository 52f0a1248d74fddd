"""Small helpers with the names of reference ``index/utils.py``."""
import datetime
import os

_COLORS = ["black", "red", "green", "yellow", "blue", "pink", "cyan", "white"]


def ensure_dir(dir_path):
    os.makedirs(dir_path, exist_ok=True)


def set_color(log, color, highlight=True):
    idx = _COLORS.index(color) if color in _COLORS else len(_COLORS) - 1
    return "\033[" + ("1;3" if highlight else "0;3") + str(idx) + "m" + log + "\033[0m"


def get_local_time():
    return datetime.datetime.now().strftime("%b-%d-%Y_%H-%M-%S")


def delete_file(filename):
    if os.path.exists(filename):
        os.remove(filename)
