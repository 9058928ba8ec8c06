from .ari_eval import ARIEvaluator, make_evaluator  # noqa: F401
