from . import stylePrediction, styleTransfer, styleTransferInferenceModel, styleTransferTrainingModel, styleLoss  # noqa: F401
