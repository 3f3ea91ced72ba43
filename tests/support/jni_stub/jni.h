/* Test-only stand-in for <jni.h>: just enough of the JNI surface for `g++ -fsyntax-only` to type-check
 * integration/jni/lzy_b200_jni.cpp in an image that has no JDK.  Signatures follow the JNI specification
 * (Java SE 17, "JNI Functions"); nothing here is ever linked or run. */
#ifndef LZYB_TEST_JNI_STUB_H
#define LZYB_TEST_JNI_STUB_H
#include <cstdint>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_OK 0
#define JNI_ABORT 2

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef jint jsize;
class _jobject {};
class _jclass : public _jobject {};
class _jstring : public _jobject {};
class _jarray : public _jobject {};
class _jbyteArray : public _jarray {};
class _jlongArray : public _jarray {};
class _jthrowable : public _jobject {};
typedef _jobject* jobject;
typedef _jclass* jclass;
typedef _jstring* jstring;
typedef _jarray* jarray;
typedef _jbyteArray* jbyteArray;
typedef _jlongArray* jlongArray;
typedef _jthrowable* jthrowable;

struct JNIEnv_ {
    const char* GetStringUTFChars(jstring, jboolean*);
    void ReleaseStringUTFChars(jstring, const char*);
    jstring NewStringUTF(const char*);
    void* GetDirectBufferAddress(jobject);
    jlong GetDirectBufferCapacity(jobject);
    jsize GetArrayLength(jarray);
    jbyteArray NewByteArray(jsize);
    void SetByteArrayRegion(jbyteArray, jsize, jsize, const jbyte*);
    void GetByteArrayRegion(jbyteArray, jsize, jsize, jbyte*);
    void SetLongArrayRegion(jlongArray, jsize, jsize, const jlong*);
    jclass FindClass(const char*);
    jint ThrowNew(jclass, const char*);
};
typedef JNIEnv_ JNIEnv;
#endif
