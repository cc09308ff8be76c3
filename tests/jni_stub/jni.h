/* Minimal stand-in for <jni.h> so that jni/rnabloom_jni.c can be syntax-checked where no JDK exists (tests/test_abi.py).
 * Declares only what the shim uses; NOT a JNI implementation. */
#ifndef RB_JNI_STUB_H
#define RB_JNI_STUB_H
#include <stdint.h>
typedef int32_t jint; typedef int64_t jlong; typedef float jfloat; typedef uint8_t jboolean;
typedef void* jobject; typedef jobject jclass; typedef jobject jstring;
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
    void* (*GetDirectBufferAddress)(JNIEnv*, jobject);
    jclass (*FindClass)(JNIEnv*, const char*);
    jint (*ThrowNew)(JNIEnv*, jclass, const char*);
    const char* (*GetStringUTFChars)(JNIEnv*, jstring, jboolean*);
    void (*ReleaseStringUTFChars)(JNIEnv*, jstring, const char*);
};
#endif
