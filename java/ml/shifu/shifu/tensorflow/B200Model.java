/*
 * B200Model - drop-in for ml.shifu.shifu.tensorflow.TensorflowModel (shifu-tensorflow-eval): the ml.shifu.shifu.core.Computable
 * contract (init / compute / releaseResource) with the TF-Java session replaced by libshifu_b200.so through the JNI shim
 * shifu-tensorflow_b200/csrc/jni_shim.c.  What callers can observe is kept: the exception type and message of every
 * configuration error (TensorflowModel.java:117-166), "TF model not initialized." before init (:55-57), idempotent init
 * (:114-116), one output only (:137-139).
 *
 * NOT compiled in the build container (no JDK there); INTEGRATION.md has the build line.  Native methods map 1:1 onto the
 * C-ABI of include/shifu_b200.h:
 *   nativeLoad       -> sb_model_load          (SavedModelBundle.load,  TensorflowModel.java:169)
 *   nativeScoreRow   -> sb_model_score_row_f64 (compute: feed/fetch/run, TensorflowModel.java:53-94)
 *   nativeScoreBatch -> sb_model_score         (new: rows scored in one call)
 *   nativeDestroy    -> sb_model_destroy
 *
 * Extra named inputs.  The reference feeds properties[inputNames[i]] for i >= 1 as constant tensors (:73-83; in its test a
 * Keras learning-phase bool, TensorflowModelTest.java:44-47).  The loader here walks the INFERENCE branch of the graph, so
 * such an input can only be honoured when it selects that branch: Boolean.FALSE or a numeric zero is accepted (and not fed),
 * a missing value is skipped exactly like the reference's catch block does (:78-80), Boolean.TRUE / non-zero asks for the
 * training branch (dropout active) and any other type cannot be a phase switch - both are rejected at init.
 */
package ml.shifu.shifu.tensorflow;

import java.util.Collections;
import java.util.List;
import java.util.Map;

import org.encog.ml.data.MLData;

import ml.shifu.shifu.container.obj.GenericModelConfig;
import ml.shifu.shifu.core.Computable;

public class B200Model implements Computable {

    static {
        System.loadLibrary("shifu_b200_jni"); // links libshifu_b200.so
    }

    /** What init() extracts from a GenericModelConfig; immutable once built. */
    private static final class Wiring {
        final String modelDir, feed, fetch, tag;

        Wiring(String modelDir, String feed, String fetch, String tag) {
            this.modelDir = modelDir;
            this.feed = feed;
            this.fetch = fetch;
            this.tag = tag;
        }
    }

    public Map<String, Object> properties = Collections.emptyMap();

    private volatile long handle = 0L; // 0 = not initialised (or released)
    private Wiring wiring;

    private static native long nativeLoad(String modelDir, String inputName, String outputName, String tag, int device,
            int precision);

    private static native double nativeScoreRow(long handle, double[] row);

    private static native float[] nativeScoreBatch(long handle, float[] rowsRowMajor, int nRows);

    private static native void nativeDestroy(long handle);

    private long live() {
        long h = handle;
        if(h == 0L) {
            throw new IllegalStateException("TF model not initialized.");
        }
        return h;
    }

    @Override
    public double compute(MLData input) {
        return nativeScoreRow(live(), input.getData()); // the double -> float cast of :64-68 happens behind the C-ABI
    }

    /** New: the only way to reach the GPU's throughput - one JNI crossing for nRows rows. */
    public float[] computeBatch(float[] rowsRowMajor, int nRows) {
        return nativeScoreBatch(live(), rowsRowMajor, nRows);
    }

    private static void need(boolean present, String what) {
        if(!present) {
            throw new RuntimeException(what + " is null"); // messages of TensorflowModel.java:117-166
        }
    }

    private static String onlyOutput(Object declared) {
        if(declared instanceof String[]) {
            String[] all = (String[]) declared;
            if(all.length != 1) {
                throw new IllegalArgumentException("Output now only support single output in inference.");
            }
            return all[0];
        }
        return declared instanceof String ? (String) declared : null;
    }

    /** rule for inputNames[1..], see the header */
    private static void checkPhaseSwitch(String name, Object value) {
        if(value == null) {
            return; // the reference logs "Invalid input" and does not feed it
        }
        boolean inference;
        if(value instanceof Boolean) {
            inference = !((Boolean) value);
        } else if(value instanceof Number) {
            inference = ((Number) value).doubleValue() == 0.0;
        } else {
            throw new IllegalArgumentException("Input " + name + " has unsupported type " + value.getClass().getName()
                    + ": only boolean / numeric inference-phase switches can be honoured.");
        }
        if(!inference) {
            throw new IllegalArgumentException("Input " + name + " = " + value + " selects the training branch of the graph; "
                    + "only inference (false / 0) is supported.");
        }
    }

    private static Wiring read(GenericModelConfig config, Map<String, Object> props) {
        List<String> feeds = config.getInputnames();
        String fetch = onlyOutput(props.get("outputnames"));
        @SuppressWarnings("unchecked")
        List<String> tags = (List<String>) props.get("tags");
        String dir = (String) props.get("modelpath");
        need(dir != null && !dir.isEmpty(), "Model path");
        need(feeds != null && !feeds.isEmpty(), "Input names");
        need(fetch != null && !fetch.isEmpty(), "Output names");
        need(tags != null && !tags.isEmpty(), "Tags");
        for(String extra: feeds.subList(1, feeds.size())) {
            checkPhaseSwitch(extra, props.get(extra));
        }
        return new Wiring(dir, feeds.get(0), fetch, tags.get(0));
    }

    /** shifu.b200.precision: fp32 (0, CUDA cores) | bf16 (1) | fp32_tc (2, default: fp32-class accuracy on tensor cores) | bf16x2 (3) */
    private static int precision() {
        String p = System.getProperty("shifu.b200.precision", "fp32_tc").toLowerCase();
        switch (p) {
            case "fp32": case "0": return 0;
            case "bf16": case "1": return 1;
            case "bf16x2": case "3": return 3;
            default: return 2;
        }
    }

    @Override
    public synchronized void init(GenericModelConfig config) {
        if(handle != 0L) {
            return;
        }
        need(config != null, "Config");
        Map<String, Object> props = config.getProperties();
        need(props != null && !props.isEmpty(), "Properties");
        this.properties = props;
        this.wiring = read(config, props);
        this.handle = nativeLoad(wiring.modelDir, wiring.feed, wiring.fetch, wiring.tag,
                Integer.getInteger("shifu.b200.device", 0), precision());
    }

    @Override
    public synchronized void releaseResource() {
        long h = handle;
        handle = 0L;
        if(h != 0L) {
            nativeDestroy(h);
        }
    }
}
