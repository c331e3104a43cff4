"""Minimal SB3 `Logger` surface used by imitation.util.logger (names + bookkeeping)."""
import collections
import os
import sys


class KVWriter:
    def write(self, key_values, key_excluded, step=0):
        raise NotImplementedError

    def close(self):
        pass


class SeqWriter:
    def write_sequence(self, sequence):
        raise NotImplementedError


class HumanOutputFormat(KVWriter, SeqWriter):
    def __init__(self, filename_or_file, max_length=36):
        self.max_length = max_length
        self.file = filename_or_file if hasattr(filename_or_file, "write") else open(filename_or_file, "w")

    def write(self, key_values, key_excluded, step=0):
        pass

    def write_sequence(self, sequence):
        pass


class _NullFormat(KVWriter):
    def write(self, key_values, key_excluded, step=0):
        pass


def make_output_format(_format, log_dir, log_suffix=""):
    os.makedirs(log_dir, exist_ok=True)
    if _format == "stdout":
        return HumanOutputFormat(sys.stdout)
    return _NullFormat()


class Logger:
    def __init__(self, folder, output_formats):
        self.name_to_value = collections.defaultdict(float)
        self.name_to_count = collections.defaultdict(int)
        self.name_to_excluded = {}
        self.level = 20
        self.dir = folder
        self.output_formats = output_formats

    def record(self, key, value, exclude=None):
        self.name_to_value[key] = value
        self.name_to_excluded[key] = exclude

    def record_mean(self, key, value, exclude=None):
        if value is None:
            return
        old_val, count = self.name_to_value[key], self.name_to_count[key]
        self.name_to_value[key] = old_val * count / (count + 1) + value / (count + 1)
        self.name_to_count[key] = count + 1
        self.name_to_excluded[key] = exclude

    def dump(self, step=0):
        for fmt in self.output_formats:
            if isinstance(fmt, KVWriter):
                fmt.write(self.name_to_value, self.name_to_excluded, step)
        self.name_to_value.clear()
        self.name_to_count.clear()
        self.name_to_excluded.clear()

    def log(self, *args, level=20):
        pass

    def debug(self, *args):
        pass

    def info(self, *args):
        pass

    def warn(self, *args):
        pass

    def error(self, *args):
        pass

    def set_level(self, level):
        self.level = level

    def get_dir(self):
        return self.dir

    def close(self):
        for fmt in self.output_formats:
            fmt.close()
