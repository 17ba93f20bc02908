"""Elapsed-time and append-to-file helpers with the reference's names (Network/utility.py:9-26)."""
import time


def calculate_time_elapsed(start):
    secs = time.time() - start
    hrs = int(secs // 3600)
    mins = int((secs % 3600) // 60)
    return hrs, mins, int(secs % 60)


def log_to_file(logfile, message):
    with open(logfile, "a") as f:
        f.write(message)
