/*
 * B200Model - drop-in for ml.shifu.shifu.tensorflow.TensorflowModel (shifu-tensorflow-eval): the same
 * ml.shifu.shifu.core.Computable contract (init / compute / releaseResource) and the same exceptions, with the
 * TF-Java session replaced by libshifu_b200.so through the JNI shim in shifu-tensorflow_b200/csrc/jni_shim.c.
 *
 * NOT compiled in the build container (no JDK there); see INTEGRATION.md for the build line.  Each native method
 * maps 1:1 onto a C-ABI entry point of include/shifu_b200.h:
 *   nativeLoad       -> sb_model_load          (SavedModelBundle.load,  TensorflowModel.java:169)
 *   nativeScoreRow   -> sb_model_score_row_f64 (compute: feed/fetch/run, TensorflowModel.java:53-94)
 *   nativeScoreBatch -> sb_model_score         (new: rows scored in one call)
 *   nativeDestroy    -> sb_model_destroy
 */
package ml.shifu.shifu.tensorflow;

import java.util.HashMap;
import java.util.List;
import java.util.Map;

import org.encog.ml.data.MLData;

import ml.shifu.shifu.container.obj.GenericModelConfig;
import ml.shifu.shifu.core.Computable;

public class B200Model implements Computable {

    static {
        System.loadLibrary("shifu_b200_jni"); // links libshifu_b200.so
    }

    public Map<String, Object> properties = new HashMap<String, Object>();

    private boolean initiate = false;
    private long handle = 0L;
    private String modelPath;
    private String[] tags;
    private String[] inputNames;
    private String outputNames;

    private static native long nativeLoad(String modelDir, String inputName, String outputName, String tag, int device,
            int precision);

    private static native double nativeScoreRow(long handle, double[] row);

    private static native float[] nativeScoreBatch(long handle, float[] rowsRowMajor, int nRows);

    private static native void nativeDestroy(long handle);

    @Override
    public double compute(MLData input) {
        if(!initiate || handle == 0L) {
            throw new IllegalStateException("TF model not initialized.");
        }
        return nativeScoreRow(handle, input.getData()); // the double -> float cast happens in the C-ABI
    }

    /** New: the only way to reach the GPU's throughput - one JNI crossing for nRows rows. */
    public float[] computeBatch(float[] rowsRowMajor, int nRows) {
        if(!initiate || handle == 0L) {
            throw new IllegalStateException("TF model not initialized.");
        }
        return nativeScoreBatch(handle, rowsRowMajor, nRows);
    }

    @Override
    @SuppressWarnings("unchecked")
    public void init(GenericModelConfig config) {
        if(this.initiate) {
            return;
        }
        if(config == null) {
            throw new RuntimeException("Config is null");
        }
        properties = config.getProperties();
        if(properties == null || properties.size() == 0) {
            throw new RuntimeException("Properties is null");
        }
        this.modelPath = (String) properties.get("modelpath");
        this.inputNames = config.getInputnames().toArray(new String[0]);
        Object outputNames = properties.get("outputnames");
        if(outputNames instanceof String) {
            this.outputNames = (String) outputNames;
        } else if(outputNames instanceof String[]) {
            String[] outputs = (String[]) outputNames;
            if(outputs.length == 1) {
                this.outputNames = outputs[0];
            } else {
                throw new IllegalArgumentException("Output now only support single output in inference.");
            }
        }
        List<String> tagList = (List<String>) properties.get("tags");
        this.tags = tagList == null ? null : tagList.toArray(new String[tagList.size()]);
        if(this.modelPath == null || this.modelPath.isEmpty()) {
            throw new RuntimeException("Model path is null");
        }
        if(this.inputNames == null || this.inputNames.length == 0) {
            throw new RuntimeException("Input names is null");
        }
        if(this.outputNames == null || this.outputNames.isEmpty()) {
            throw new RuntimeException("Output names is null");
        }
        if(this.tags == null || this.tags.length == 0) {
            throw new RuntimeException("Tags is null");
        }
        int device = Integer.getInteger("shifu.b200.device", 0);
        int precision = Integer.getInteger("shifu.b200.precision", 0); // 0 = fp32 parity mode, 1 = bf16
        this.handle = nativeLoad(modelPath, inputNames[0], this.outputNames, tags[0], device, precision);
        initiate = true;
    }

    @Override
    public void releaseResource() {
        if(handle != 0L) {
            nativeDestroy(handle);
            handle = 0L;
        }
        initiate = false;
    }
}
