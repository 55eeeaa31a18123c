def subplot(*args, **kwargs):
    raise NotImplementedError("dg_util shim: drawing is visualisation only")


def draw_contrast_text_cv2(*args, **kwargs):
    raise NotImplementedError("dg_util shim: drawing is visualisation only")
