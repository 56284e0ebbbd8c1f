// Minimal stand-in for the JDK's <jni.h>, written for this repo (no JDK exists in the build image).  Only what jni_shim.cpp uses:
// the primitive typedefs and a JNIEnv whose member functions forward through the interface function table at the slot indices the
// JNI specification fixes ("Interface Function Table", JNI spec chapter 4; the same positions every JVM implements):
//   6 FindClass, 14 ThrowNew, 171 GetArrayLength, 212 SetLongArrayRegion, 222 GetPrimitiveArrayCritical,
//   223 ReleasePrimitiveArrayCritical, 228 ExceptionCheck.
// A production build compiles jni_shim.cpp against the JDK's own header instead (-DFILO_USE_SYSTEM_JNI -I$JAVA_HOME/include ...);
// the exported symbols and their signatures are identical.  tests/cpp/host_mirror_gpu.cpp fills a table with host functions and
// drives the shim through it.
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef double jdouble;
typedef jint jsize;
struct _jobject;
typedef struct _jobject* jobject;
typedef jobject jclass;
typedef jobject jarray;
typedef jarray jintArray;
typedef jarray jlongArray;
typedef jarray jdoubleArray;
#define JNI_ABORT 2
#define JNI_FALSE 0
#define JNI_TRUE 1
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define FILO_JNI_SLOTS 233
struct JNINativeInterface_ { void* slot[FILO_JNI_SLOTS]; };
#ifdef __cplusplus
}
struct JNIEnv_ {
  const struct JNINativeInterface_* functions;
  jclass FindClass(const char* name) { return reinterpret_cast<jclass (*)(JNIEnv_*, const char*)>(functions->slot[6])(this, name); }
  jint ThrowNew(jclass c, const char* msg) { return reinterpret_cast<jint (*)(JNIEnv_*, jclass, const char*)>(functions->slot[14])(this, c, msg); }
  jsize GetArrayLength(jarray a) { return reinterpret_cast<jsize (*)(JNIEnv_*, jarray)>(functions->slot[171])(this, a); }
  void SetLongArrayRegion(jlongArray a, jsize start, jsize len, const jlong* buf) {
    reinterpret_cast<void (*)(JNIEnv_*, jlongArray, jsize, jsize, const jlong*)>(functions->slot[212])(this, a, start, len, buf);
  }
  void* GetPrimitiveArrayCritical(jarray a, jboolean* isCopy) { return reinterpret_cast<void* (*)(JNIEnv_*, jarray, jboolean*)>(functions->slot[222])(this, a, isCopy); }
  void ReleasePrimitiveArrayCritical(jarray a, void* carray, jint mode) {
    reinterpret_cast<void (*)(JNIEnv_*, jarray, void*, jint)>(functions->slot[223])(this, a, carray, mode);
  }
  jboolean ExceptionCheck() { return reinterpret_cast<jboolean (*)(JNIEnv_*)>(functions->slot[228])(this); }
};
typedef JNIEnv_ JNIEnv;
#else
typedef const struct JNINativeInterface_* JNIEnv;
#endif
