"""Import stub so the reference's utils.py (which annotates with pyscipopt types) can be imported for its
``load_batch``; SCIP itself is absent.  Test infrastructure only."""
from . import scip  # noqa: F401


class SCIP_RESULT:  # model_benchmarker.py:34, 157
    SUCCESS = "SUCCESS"
