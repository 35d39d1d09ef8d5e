class Model:  # annotation targets only (utils.py:35)
    pass


class Row:
    pass


class Cutsel:  # base class of the reference's CustomCutsel (model_benchmarker.py:36, 41); SCIP sets ``model``
    model = None
