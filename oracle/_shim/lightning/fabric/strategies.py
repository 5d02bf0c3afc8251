class DeepSpeedStrategy:  # placeholder: isinstance checks only
    pass


class FSDPStrategy:
    pass
