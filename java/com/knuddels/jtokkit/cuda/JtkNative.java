package com.knuddels.jtokkit.cuda;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/**
 * Panama FFM (JDK 22+) binding of include/jtokkit_b200.h.  NOT COMPILED IN THIS REPOSITORY'S ENVIRONMENT (no JDK in the image);
 * it documents the exact downcall signatures a JTokkit maintainer adds.  No JNI glue and no C code are needed.
 */
final class JtkNative {
	private static final Linker LINKER = Linker.nativeLinker();
	private static final SymbolLookup LIB = SymbolLookup.libraryLookup(System.getProperty("jtokkit.b200.lib", "libjtokkit_b200.so"), Arena.global());

	private static MethodHandle fn(final String name, final FunctionDescriptor fd) {
		return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), fd);
	}

	/** int jtk_encoding_create(const jtk_params*, const int* devices, int ndev, jtk_encoding** out) */
	static final MethodHandle ENCODING_CREATE = fn("jtk_encoding_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS));
	/** void jtk_encoding_destroy(jtk_encoding*) */
	static final MethodHandle ENCODING_DESTROY = fn("jtk_encoding_destroy", FunctionDescriptor.ofVoid(ADDRESS));
	/** int jtk_encode_batch(jtk_encoding*, const uint8_t* utf8, const int64_t* doc_off, int64_t ndocs, uint32_t flags, jtk_result** out) */
	static final MethodHandle ENCODE_BATCH = fn("jtk_encode_batch", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, JAVA_INT, ADDRESS));
	/** int jtk_encode_batch_special(...): same signature; special tokens in the text become their ids (not in the reference) */
	static final MethodHandle ENCODE_BATCH_SPECIAL = fn("jtk_encode_batch_special", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, JAVA_INT, ADDRESS));
	static final MethodHandle RESULT_NUM_TOKENS = fn("jtk_result_num_tokens", FunctionDescriptor.of(JAVA_LONG, ADDRESS));
	static final MethodHandle RESULT_IDS = fn("jtk_result_ids", FunctionDescriptor.of(ADDRESS, ADDRESS));
	static final MethodHandle RESULT_TOKEN_OFFSETS = fn("jtk_result_token_offsets", FunctionDescriptor.of(ADDRESS, ADDRESS));
	static final MethodHandle RESULT_DOC_STATUS = fn("jtk_result_doc_status", FunctionDescriptor.of(ADDRESS, ADDRESS));
	static final MethodHandle RESULT_DEVICE_MS = fn("jtk_result_device_ms", FunctionDescriptor.of(JAVA_DOUBLE, ADDRESS));
	static final MethodHandle RESULT_FREE = fn("jtk_result_free", FunctionDescriptor.ofVoid(ADDRESS));
	/** int jtk_decode_batch(jtk_encoding*, const int32_t* ids, const int64_t* tok_off, int64_t ndocs, jtk_result** out) */
	static final MethodHandle DECODE_BATCH = fn("jtk_decode_batch", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS));
	static final MethodHandle RESULT_BYTES = fn("jtk_result_bytes", FunctionDescriptor.of(ADDRESS, ADDRESS));
	static final MethodHandle RESULT_BYTE_OFFSETS = fn("jtk_result_byte_offsets", FunctionDescriptor.of(ADDRESS, ADDRESS));
	static final MethodHandle RESULT_BAD_IDS = fn("jtk_result_bad_ids", FunctionDescriptor.of(ADDRESS, ADDRESS));
	/** void* jtk_host_alloc(int64_t) / void jtk_host_free(void*): pinned staging for large batches */
	static final MethodHandle HOST_ALLOC = fn("jtk_host_alloc", FunctionDescriptor.of(ADDRESS, JAVA_LONG));
	static final MethodHandle HOST_FREE = fn("jtk_host_free", FunctionDescriptor.ofVoid(ADDRESS));
	static final MethodHandle LAST_ERROR = fn("jtk_last_error", FunctionDescriptor.of(ADDRESS));

	static final int JTK_OK = 0, JTK_E_PATTERN_UNSUPPORTED = -3;
	static final int CHECK_SPECIAL = 1, COUNT_ONLY = 2;
	static final int DOC_HAS_SPECIAL = 1, DOC_UNKNOWN_BYTES = 2, DOC_UNKNOWN_ID = 4, DOC_PATTERN_STACK = 8;

	static String lastError() throws Throwable {
		return ((MemorySegment) LAST_ERROR.invokeExact()).reinterpret(4096).getString(0);
	}

	private JtkNative() {
	}
}
