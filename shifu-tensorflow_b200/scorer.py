"""Host-side mirror of the reference Java scorer `ml.shifu.shifu.tensorflow.TensorflowModel`
(shifu-tensorflow-eval/src/main/java/ml/shifu/shifu/tensorflow/TensorflowModel.java), an implementation of Shifu's
`Computable` (init / compute / releaseResource), with the same error behaviour, plus the batched entry point the
per-row Session.run of the reference never had.  The Java class of the same shape that binds the same C-ABI through
JNI is java/ml/shifu/shifu/tensorflow/B200Model.java.

    config = {"inputnames": ["shifu_input_0"],
              "properties": {"modelpath": "/path/to/saved_model_dir", "outputnames": "shifu_output_0",
                             "tags": ["serve"], "algorithm": "tensorflow", "normtype": "ZSCALE"}}
    m = TensorflowModel(); m.init(config); score = m.compute(row_of_doubles)
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import _capi as capi


class IllegalStateException(RuntimeError):
    """TensorflowModel.compute before init (TensorflowModel.java:55-57)."""


class IllegalArgumentException(ValueError):
    """more than one output name (TensorflowModel.java:137-139)."""


def check_phase_switch(name: str, value: Any) -> None:
    """rule for GenericModelConfig inputnames[1:] (see TensorflowModel.init below and java/.../B200Model.java)"""
    if value is None:
        return
    if isinstance(value, (bool, int, float, np.bool_, np.integer, np.floating)):
        if bool(value):
            raise IllegalArgumentException("Input %s = %r selects the training branch of the graph; only inference "
                                           "(false / 0) is supported." % (name, value))
        return
    raise IllegalArgumentException("Input %s has unsupported type %s: only boolean / numeric inference-phase switches can be "
                                   "honoured." % (name, type(value).__name__))


class TensorflowModel:
    def __init__(self, device: int = 0, precision: int = capi.PREC_FP32):
        self.properties: Dict[str, Any] = {}
        self.initiate = False
        self.modelPath: Optional[str] = None
        self._model: Optional[capi.Model] = None
        self.config = None
        self.tags: Optional[List[str]] = None
        self.inputNames: Optional[List[str]] = None
        self.outputNames: Optional[str] = None
        self._device, self._precision = device, precision

    # -- Computable.init (TensorflowModel.java:112-172) --
    def init(self, config) -> None:
        if self.initiate:                                   # idempotent (:114-116)
            return
        if config is None:
            raise RuntimeError("Config is null")
        self.config = config
        get = (lambda k: config.get(k)) if isinstance(config, dict) else (lambda k: getattr(config, k, None))
        self.properties = get("properties")
        if self.properties is None or len(self.properties) == 0:
            raise RuntimeError("Properties is null")
        self.modelPath = self.properties.get("modelpath")
        names = get("inputnames")
        self.inputNames = list(names) if names is not None else None
        output_names = self.properties.get("outputnames")
        if isinstance(output_names, str):
            self.outputNames = output_names
        elif isinstance(output_names, (list, tuple)):
            if len(output_names) == 1:
                self.outputNames = output_names[0]
            else:
                raise IllegalArgumentException("Output now only support single output in inference.")
        tag_list = self.properties.get("tags")
        self.tags = list(tag_list) if tag_list is not None else None
        if not self.modelPath:
            raise RuntimeError("Model path is null")
        if not self.inputNames:
            raise RuntimeError("Input names is null")
        if not self.outputNames:
            raise RuntimeError("Output names is null")
        if not self.tags:
            raise RuntimeError("Tags is null")
        # SavedModelBundle.load(modelPath, tags) + feed inputNames[0] / fetch outputNames by op name (:71,85,169).
        # Extra named inputs (inputNames[1:], fed from `properties` as constants, :73-83 - in the reference's test a Keras
        # learning-phase bool): the loader walks the INFERENCE branch of the graph, so such an input is honoured only when it
        # selects that branch.  False / 0 -> accepted (not fed); missing -> skipped like the reference's catch block (:78-80);
        # True / non-zero (training branch: dropout active) or any other type -> rejected here, at init.
        for name in self.inputNames[1:]:
            check_phase_switch(name, self.properties.get(name))
        self._model = capi.Model.load(self.modelPath, self.inputNames[0], self.outputNames, tag=self.tags[0],
                                      device=self._device, precision=self._precision)
        self.initiate = True

    # -- Computable.compute (TensorflowModel.java:53-94): double[] -> float[] -> [1,n] forward -> double --
    def compute(self, input) -> float:
        if not self.initiate or self._model is None:
            raise IllegalStateException("TF model not initialized.")
        data = input.getData() if hasattr(input, "getData") else input
        return self._model.score_row_f64(np.asarray(data, dtype=np.float64))

    # -- new: the whole table in one call (rows are scored in parallel on the GPU) --
    def computeBatch(self, rows) -> np.ndarray:
        if not self.initiate or self._model is None:
            raise IllegalStateException("TF model not initialized.")
        X = np.asarray(rows, dtype=np.float64).astype(np.float32)     # the same double -> float cast, vectorised
        return self._model.score(X).astype(np.float64)

    def releaseResource(self) -> None:
        """The reference never closes its bundle (TensorflowModel.java:175-176); here device memory is returned."""
        if self._model is not None:
            self._model.close()
            self._model = None
        self.initiate = False
