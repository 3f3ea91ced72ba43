"""The reference-side binding of INTEGRATION.md §1 exists as sources under integration/.  This image has no JDK, so
the check is what can be done without one: the JNI shim type-checks against the real C ABI headers (and a stand-in
jni.h), and the Java `native` declarations and the shim's exported functions are the same set."""
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SHIM = ROOT / "integration" / "jni" / "lzy_b200_jni.cpp"
JAVA = ROOT / "integration" / "java" / "ai" / "lzy" / "slots" / "b200"


def test_jni_shim_type_checks_against_the_c_abi():
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Werror",
                        f"-I{ROOT / 'tests' / 'support' / 'jni_stub'}", f"-I{ROOT / 'include'}", str(SHIM)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_native_declarations_and_shim_exports_are_the_same_set():
    declared = set(re.findall(r"public static native [\w\[\]]+ (\w+)\(", (JAVA / "Native.java").read_text()))
    exported = set(re.findall(r"Java_ai_lzy_slots_b200_Native_(\w+)\(", SHIM.read_text()))
    assert declared and declared == exported


def test_shim_only_calls_functions_the_library_exports():
    from lzy_b200 import _native
    called = set(re.findall(r"\b(lzyb_\w+)\(", SHIM.read_text()))
    assert called <= set(_native.EXPORTED_SYMBOLS), called - set(_native.EXPORTED_SYMBOLS)


def test_java_classes_implement_the_reference_seams():
    text = {p.name: p.read_text() for p in JAVA.glob("*.java")}
    assert "implements OutputSlotBackend" in text["HbmOutputBackend.java"]
    assert "implements InputSlotBackend" in text["HbmInputBackend.java"]
    assert "implements InputTransfer" in text["NvlinkInputTransfer.java"]
    for name, methods in (("HbmOutputBackend.java", ("waitCompleted()", "readFromOffset(long", "close()")),
                          ("HbmInputBackend.java", ("openChannel()", "toOutput()", "close()")),
                          ("NvlinkInputTransfer.java", ("transferChunkTo(SeekableByteChannel",))):
        for m in methods:
            assert m in text[name], (name, m)
