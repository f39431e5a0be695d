/* TEST INFRASTRUCTURE: a minimal stand-in for the JDK's <jni.h>, just enough of the JNI 1.6 C
 * interface to type-check maelstrom_b200/csrc/ms_jni.c on a box without a JDK
 * (tests/test_boundary_files.py).  Names, signatures and calling convention macros follow the
 * JNI specification; nothing here can run. */
#ifndef MS_JNI_STUB_H
#define MS_JNI_STUB_H
#include <stdint.h>
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
typedef int32_t jint;
typedef int64_t jlong;
typedef double jdouble;
typedef uint8_t jboolean;
typedef struct _jobject* jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef struct _jmethodID* jmethodID;
struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
  jclass (*GetObjectClass)(JNIEnv*, jobject);
  jmethodID (*GetMethodID)(JNIEnv*, jclass, const char*, const char*);
  jint (*CallIntMethod)(JNIEnv*, jobject, jmethodID, ...);
  jstring (*NewStringUTF)(JNIEnv*, const char*);
  const char* (*GetStringUTFChars)(JNIEnv*, jstring, jboolean*);
  void (*ReleaseStringUTFChars)(JNIEnv*, jstring, const char*);
  jobject (*NewDirectByteBuffer)(JNIEnv*, void*, jlong);
  void* (*GetDirectBufferAddress)(JNIEnv*, jobject);
  void (*DeleteLocalRef)(JNIEnv*, jobject);
  jboolean (*ExceptionCheck)(JNIEnv*);
};
#endif
