class Model:  # annotation targets only (utils.py:35)
    pass


class Row:
    pass
