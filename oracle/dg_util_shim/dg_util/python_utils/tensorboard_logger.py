"""Import-level stand-in for dg_util.python_utils.tensorboard_logger (TEST INFRASTRUCTURE ONLY): logging is a no-op."""


class Logger(object):
    def __init__(self, *args, **kwargs):
        pass

    def dict_log(self, *args, **kwargs):
        pass

    def image_summary(self, *args, **kwargs):
        pass

    def network_conv_summary(self, *args, **kwargs):
        pass
