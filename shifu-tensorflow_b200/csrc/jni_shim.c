/*
 * JNI shim: pure marshalling between ml.shifu.shifu.tensorflow.B200Model (java/) and the C-ABI in
 * include/shifu_b200.h.  Compiled only where a JDK is present:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude \
 *       shifu-tensorflow_b200/csrc/jni_shim.c -Lshifu-tensorflow_b200/lib -lshifu_b200 -o libshifu_b200_jni.so
 * Errors of the C-ABI become java.lang.RuntimeException carrying sb_last_error().
 */
#if defined(__has_include)
#if __has_include(<jni.h>)
#define SB_HAVE_JNI 1
#endif
#endif

#ifdef SB_HAVE_JNI
#include <jni.h>
#include <stdint.h>
#include <stdlib.h>
#include "../../include/shifu_b200.h"

static void sb_throw(JNIEnv* env, const char* cls) {
  jclass c = (*env)->FindClass(env, cls);
  if (c) (*env)->ThrowNew(env, c, sb_last_error());
}

JNIEXPORT jlong JNICALL Java_ml_shifu_shifu_tensorflow_B200Model_nativeLoad(JNIEnv* env, jclass k, jstring dir, jstring in,
                                                                           jstring out, jstring tag, jint device, jint prec) {
  const char* d = (*env)->GetStringUTFChars(env, dir, 0);
  const char* i = (*env)->GetStringUTFChars(env, in, 0);
  const char* o = (*env)->GetStringUTFChars(env, out, 0);
  const char* t = (*env)->GetStringUTFChars(env, tag, 0);
  sb_model_t* m = NULL;
  int s = sb_model_load(d, i, o, t, device, prec, &m);
  (*env)->ReleaseStringUTFChars(env, dir, d);
  (*env)->ReleaseStringUTFChars(env, in, i);
  (*env)->ReleaseStringUTFChars(env, out, o);
  (*env)->ReleaseStringUTFChars(env, tag, t);
  if (s != SB_OK) { sb_throw(env, "java/lang/RuntimeException"); return 0; }
  return (jlong)(intptr_t)m;
}

JNIEXPORT jdouble JNICALL Java_ml_shifu_shifu_tensorflow_B200Model_nativeScoreRow(JNIEnv* env, jclass k, jlong h, jdoubleArray row) {
  jsize n = (*env)->GetArrayLength(env, row);
  jdouble* p = (*env)->GetDoubleArrayElements(env, row, 0);
  double r = 0.0;
  int s = sb_model_score_row_f64((sb_model_t*)(intptr_t)h, p, (int32_t)n, &r);
  (*env)->ReleaseDoubleArrayElements(env, row, p, JNI_ABORT);
  if (s != SB_OK) sb_throw(env, s == SB_ERR_STATE ? "java/lang/IllegalStateException" : "java/lang/RuntimeException");
  return r;
}

JNIEXPORT jfloatArray JNICALL Java_ml_shifu_shifu_tensorflow_B200Model_nativeScoreBatch(JNIEnv* env, jclass k, jlong h,
                                                                                      jfloatArray rows, jint n_rows) {
  jfloat* p = (*env)->GetFloatArrayElements(env, rows, 0);
  jfloatArray out = (*env)->NewFloatArray(env, n_rows);
  float* o = (float*)malloc(sizeof(float) * (size_t)(n_rows > 0 ? n_rows : 1));
  int s = sb_model_score((sb_model_t*)(intptr_t)h, p, (int64_t)n_rows, o);
  (*env)->ReleaseFloatArrayElements(env, rows, p, JNI_ABORT);
  if (s == SB_OK) (*env)->SetFloatArrayRegion(env, out, 0, n_rows, o);
  free(o);
  if (s != SB_OK) { sb_throw(env, s == SB_ERR_STATE ? "java/lang/IllegalStateException" : "java/lang/RuntimeException"); return NULL; }
  return out;
}

JNIEXPORT void JNICALL Java_ml_shifu_shifu_tensorflow_B200Model_nativeDestroy(JNIEnv* env, jclass k, jlong h) {
  sb_model_destroy((sb_model_t*)(intptr_t)h);
}
#else
/* no JDK on this machine: nothing to build (the ctypes binding in _capi.py covers the same C-ABI) */
typedef int sb_jni_shim_not_built;
#endif
