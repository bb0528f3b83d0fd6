class AvgPooling:  # noqa: D401 - import-only
    pass


class MaxPooling:
    pass
