from .ari_eval import ARIEvaluator, make_evaluator  # noqa: F401
from .loop import evaluate  # noqa: F401
