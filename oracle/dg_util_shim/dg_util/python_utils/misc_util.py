import time


def get_time_str():
    # constants.py:23 only needs a unique-ish string for the checkpoint dir.
    return time.strftime("%Y_%m_%d_%H_%M_%S")


def resize(image, size, *args, **kwargs):  # datasets only; never hit by the oracle
    raise NotImplementedError("dg_util shim: misc_util.resize is data-pipeline only")
