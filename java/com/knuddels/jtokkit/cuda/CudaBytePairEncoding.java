package com.knuddels.jtokkit.cuda;

import com.knuddels.jtokkit.api.Encoding;
import com.knuddels.jtokkit.api.EncodingResult;
import com.knuddels.jtokkit.api.GptBytePairEncodingParams;

import java.lang.foreign.Arena;
import java.lang.foreign.MemoryLayout;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.StructLayout;
import java.nio.charset.StandardCharsets;
import java.util.ArrayList;
import java.util.Collections;
import java.util.List;
import java.util.Map;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/**
 * Drop-in replacement for com.knuddels.jtokkit.GptBytePairEncoding (same Encoding contract, same exceptions) that forwards to
 * libjtokkit_b200.so.  NOT COMPILED HERE (no JDK in the image).  EncodingFactory.fromParameters (EncodingFactory.java:117-119)
 * becomes {@code return new CudaBytePairEncoding(parameters);}.
 */
public final class CudaBytePairEncoding implements Encoding, AutoCloseable {
	private static final StructLayout PARAMS = MemoryLayout.structLayout(
			ADDRESS.withName("name"), ADDRESS.withName("pattern"), JAVA_INT.withName("pattern_flags"), MemoryLayout.paddingLayout(4),
			ADDRESS.withName("vocab_bytes"), ADDRESS.withName("vocab_off"), ADDRESS.withName("vocab_ranks"), JAVA_LONG.withName("vocab_size"),
			ADDRESS.withName("special_bytes"), ADDRESS.withName("special_off"), ADDRESS.withName("special_ids"), JAVA_LONG.withName("special_size"));

	private final String name;
	private final MemorySegment handle;
	private final boolean[] singleByteTokens = new boolean[256]; // which single bytes are tokens (for the exception text of TokenEncoder.java:67)

	/** Devices from the system property jtokkit.cuda.devices ("0,1,2,3" or "all"; default: device 0). */
	public CudaBytePairEncoding(final GptBytePairEncodingParams params) {
		this(params, devicesFromProperty());
	}

	private static int[] devicesFromProperty() {
		final String v = System.getProperty("jtokkit.cuda.devices", "0").trim();
		if (v.equals("all")) {
			final int n = Integer.getInteger("jtokkit.cuda.deviceCount", 8); // one box of B200s
			final int[] all = new int[n];
			for (int i = 0; i < n; i++) all[i] = i;
			return all;
		}
		final String[] parts = v.split(",");
		final int[] out = new int[parts.length];
		for (int i = 0; i < parts.length; i++) out[i] = Integer.parseInt(parts[i].trim());
		return out;
	}

	/**
	 * @param devices CUDA devices the tables are replicated on; a batch is cut into byte-balanced document chunks and chunk c runs on
	 *                devices[c % devices.length] inside jtk_encode_batch (AbstractMultiThreadedBenchmark.java:34-45 with GPUs as workers)
	 */
	public CudaBytePairEncoding(final GptBytePairEncodingParams params, final int[] devices) {
		this.name = params.getName();
		for (final byte[] k : params.getEncoder().keySet()) if (k.length == 1) singleByteTokens[k[0] & 0xFF] = true;
		try (Arena arena = Arena.ofConfined()) {
			final MemorySegment p = arena.allocate(PARAMS);
			p.set(ADDRESS, 0, arena.allocateFrom(params.getName()));
			p.set(ADDRESS, 8, arena.allocateFrom(params.getPattern().pattern()));
			p.set(JAVA_INT, 16, params.getPattern().flags());
			// Map<byte[], Integer> -> (bytes, offsets, ranks)
			final Map<byte[], Integer> enc = params.getEncoder();
			long total = 0;
			for (final byte[] k : enc.keySet()) total += k.length;
			final MemorySegment vb = arena.allocate(Math.max(total, 1)), vo = arena.allocate(JAVA_LONG, enc.size() + 1L), vr = arena.allocate(JAVA_INT, Math.max(enc.size(), 1));
			long pos = 0;
			int i = 0;
			for (final Map.Entry<byte[], Integer> e : enc.entrySet()) {
				vo.setAtIndex(JAVA_LONG, i, pos);
				MemorySegment.copy(e.getKey(), 0, vb, JAVA_BYTE, pos, e.getKey().length);
				vr.setAtIndex(JAVA_INT, i, e.getValue());
				pos += e.getKey().length;
				i++;
			}
			vo.setAtIndex(JAVA_LONG, i, pos);
			p.set(ADDRESS, 24, vb);
			p.set(ADDRESS, 32, vo);
			p.set(ADDRESS, 40, vr);
			p.set(JAVA_LONG, 48, enc.size());
			// Map<String, Integer> special tokens -> UTF-8 (bytes, offsets, ids)
			final Map<String, Integer> sp = params.getSpecialTokensEncoder();
			final List<byte[]> sk = new ArrayList<>();
			long stotal = 0;
			for (final String k : sp.keySet()) {
				final byte[] b = k.getBytes(StandardCharsets.UTF_8);
				sk.add(b);
				stotal += b.length;
			}
			final MemorySegment sb = arena.allocate(Math.max(stotal, 1)), so = arena.allocate(JAVA_LONG, sp.size() + 1L), si = arena.allocate(JAVA_INT, Math.max(sp.size(), 1));
			pos = 0;
			i = 0;
			for (final Map.Entry<String, Integer> e : sp.entrySet()) {
				so.setAtIndex(JAVA_LONG, i, pos);
				MemorySegment.copy(sk.get(i), 0, sb, JAVA_BYTE, pos, sk.get(i).length);
				si.setAtIndex(JAVA_INT, i, e.getValue());
				pos += sk.get(i).length;
				i++;
			}
			so.setAtIndex(JAVA_LONG, i, pos);
			p.set(ADDRESS, 56, sb);
			p.set(ADDRESS, 64, so);
			p.set(ADDRESS, 72, si);
			p.set(JAVA_LONG, 80, sp.size());
			final MemorySegment out = arena.allocate(ADDRESS);
			final MemorySegment devs = arena.allocate(JAVA_INT, Math.max(devices.length, 1));
			for (int d = 0; d < devices.length; d++) devs.setAtIndex(JAVA_INT, d, devices[d]);
			final int rc = (int) JtkNative.ENCODING_CREATE.invokeExact(p, devs, devices.length, out);
			if (rc == JtkNative.JTK_E_PATTERN_UNSUPPORTED) throw new IllegalArgumentException("split pattern not supported on the device: " + JtkNative.lastError());
			if (rc != JtkNative.JTK_OK) throw new IllegalStateException(JtkNative.lastError());
			this.handle = out.get(ADDRESS, 0);
		} catch (final RuntimeException e) {
			throw e;
		} catch (final Throwable t) {
			throw new IllegalStateException(t);
		}
	}

	/** The new batch entry point: primitive arrays, no boxing. ids of document d are ids[offsets[d] .. offsets[d+1]). */
	public static final class Batch {
		public final int[] ids;
		public final long[] tokenOffsets;
		public final int[] docStatus;

		Batch(final int[] ids, final long[] tokenOffsets, final int[] docStatus) {
			this.ids = ids;
			this.tokenOffsets = tokenOffsets;
			this.docStatus = docStatus;
		}
	}

	public Batch encodeBatch(final List<String> texts, final boolean ordinary, final boolean countOnly) {
		return encodeBatch(texts, ordinary, countOnly, false);
	}

	/** withSpecialTokens: tiktoken's allowed_special="all" (jtk_encode_batch_special); the reference has no such mode. */
	public Batch encodeBatch(final List<String> texts, final boolean ordinary, final boolean countOnly, final boolean withSpecialTokens) {
		try (Arena arena = Arena.ofConfined()) {
			// String.getBytes(UTF_8), exactly as ImmutableByteArray.from (ImmutableByteArray.java:16-19): lone surrogates become '?'.
			// Never GetStringUTFChars / modified UTF-8.
			final byte[][] utf8 = new byte[texts.size()][];
			final boolean big = utf8.length >= 4096; // transcoding and flattening are per-document work: spread them over the common pool
			if (big) {
				java.util.stream.IntStream.range(0, utf8.length).parallel()
						.forEach(d -> utf8[d] = texts.get(d) == null ? new byte[0] : texts.get(d).getBytes(StandardCharsets.UTF_8));
			} else {
				for (int d = 0; d < utf8.length; d++) utf8[d] = texts.get(d) == null ? new byte[0] : texts.get(d).getBytes(StandardCharsets.UTF_8);
			}
			long total = 0;
			for (int d = 0; d < utf8.length; d++) total += utf8[d].length;
			// Large batches are flattened into PINNED memory (jtk_host_alloc): the copy-in then runs at PCIe speed (~55 GB/s measured);
			// ordinary (pageable) native memory is copied in by the driver at ~8 GB/s.  Small batches stay in the arena.
			final boolean pinned = total >= (1L << 20);
			final MemorySegment pinnedBase = pinned ? (MemorySegment) JtkNative.HOST_ALLOC.invokeExact(total) : MemorySegment.NULL;
			final MemorySegment bytes = pinned ? pinnedBase.reinterpret(total) : arena.allocate(Math.max(total, 1), 16);
			final MemorySegment off = arena.allocate(JAVA_LONG, utf8.length + 1L);
			long pos = 0;
			final long[] starts = new long[utf8.length];
			for (int d = 0; d < utf8.length; d++) {
				off.setAtIndex(JAVA_LONG, d, pos);
				starts[d] = pos;
				pos += utf8[d].length;
			}
			if (big && pinned) { // the pinned segment has the global scope (any thread may write it); arena memory is confined to this thread
				java.util.stream.IntStream.range(0, utf8.length).parallel()
						.forEach(d -> MemorySegment.copy(utf8[d], 0, bytes, JAVA_BYTE, starts[d], utf8[d].length));
			} else {
				for (int d = 0; d < utf8.length; d++) MemorySegment.copy(utf8[d], 0, bytes, JAVA_BYTE, starts[d], utf8[d].length);
			}
			off.setAtIndex(JAVA_LONG, utf8.length, pos);
			final MemorySegment out = arena.allocate(ADDRESS);
			final int flags = (ordinary || withSpecialTokens ? 0 : JtkNative.CHECK_SPECIAL) | (countOnly ? JtkNative.COUNT_ONLY : 0);
			final int rc;
			try {
				rc = withSpecialTokens ? (int) JtkNative.ENCODE_BATCH_SPECIAL.invokeExact(handle, bytes, off, (long) utf8.length, flags, out)
				                       : (int) JtkNative.ENCODE_BATCH.invokeExact(handle, bytes, off, (long) utf8.length, flags, out);
			} finally {
				if (pinned) JtkNative.HOST_FREE.invokeExact(pinnedBase); // the call has consumed its input
			}
			if (rc != JtkNative.JTK_OK) throw new IllegalStateException(JtkNative.lastError());
			final MemorySegment r = out.get(ADDRESS, 0);
			try {
				final long n = (long) JtkNative.RESULT_NUM_TOKENS.invokeExact(r);
				final int[] ids = countOnly ? new int[0] : ((MemorySegment) JtkNative.RESULT_IDS.invokeExact(r)).reinterpret(4 * n).toArray(JAVA_INT);
				final long[] tok = ((MemorySegment) JtkNative.RESULT_TOKEN_OFFSETS.invokeExact(r)).reinterpret(8L * (utf8.length + 1)).toArray(JAVA_LONG);
				final int[] st = ((MemorySegment) JtkNative.RESULT_DOC_STATUS.invokeExact(r)).reinterpret(4L * utf8.length).toArray(JAVA_INT);
				return new Batch(ids, tok, st);
			} finally {
				JtkNative.RESULT_FREE.invokeExact(r);
			}
		} catch (final RuntimeException e) {
			throw e;
		} catch (final Throwable t) {
			throw new IllegalStateException(t);
		}
	}

	private void raise(final int status, final String text) {
		// the reference's exception types and messages (GptBytePairEncoding.java:54, TokenEncoder.java:67)
		if ((status & JtkNative.DOC_HAS_SPECIAL) != 0) throw new UnsupportedOperationException("Encoding special tokens is not supported yet.");
		if ((status & JtkNative.DOC_UNKNOWN_BYTES) != 0) {
			// TokenEncoder.java:67 appends the offending part (ImmutableByteArray.toString() = Arrays.toString of ONE byte: every part that is not
			// a token is a single byte).  The device reports the document, not the part; when only one byte value without a single-byte token
			// occurs in the text the part is known, otherwise the payload is left out rather than guessed.
			int only = -1;
			boolean unique = true;
			if (text != null)
				for (final byte b : text.getBytes(StandardCharsets.UTF_8))
					if (!singleByteTokens[b & 0xFF]) {
						if (only >= 0 && only != (b & 0xFF)) unique = false;
						only = b & 0xFF;
					}
			throw new IllegalArgumentException(only >= 0 && unique ? "Unknown token for encoding: [" + (byte) only + "]" : "Unknown token for encoding");
		}
		// general split patterns only: java.util.regex dies the same way on texts that recurse too deep
		if ((status & JtkNative.DOC_PATTERN_STACK) != 0) throw new StackOverflowError("split pattern exhausted the device backtracking stack");
	}

	private List<Integer> encodeOne(final String text, final boolean ordinary) {
		if (text == null) return Collections.emptyList(); // GptBytePairEncoding.java:48-50,72-74
		final Batch b = encodeBatch(Collections.singletonList(text), ordinary, false);
		raise(b.docStatus[0], text);
		final List<Integer> out = new ArrayList<>(b.ids.length);
		for (final int id : b.ids) out.add(id);
		return out;
	}

	@Override
	public List<Integer> encode(final String text) {
		return encodeOne(text, false);
	}

	@Override
	public List<Integer> encodeOrdinary(final String text) {
		return encodeOne(text, true);
	}

	/** New method (not part of api/Encoding.java): every registered special token in the text becomes its id. */
	public List<Integer> encodeWithSpecialTokens(final String text) {
		if (text == null) return Collections.emptyList();
		final Batch b = encodeBatch(Collections.singletonList(text), true, false, true);
		raise(b.docStatus[0], text);
		final List<Integer> out = new ArrayList<>(b.ids.length);
		for (final int id : b.ids) out.add(id);
		return out;
	}

	/** encode(text, maxTokens): full device encode, clip, then the reference's back-off loop verbatim (:90-100) with the JVM's own decoder. */
	private EncodingResult encodeMax(final String text, final int maxTokens, final boolean ordinary) {
		if (text == null) return new EncodingResult(Collections.emptyList(), false);
		final List<Integer> all = encodeOne(text, ordinary);
		final List<Integer> out = all.subList(0, Math.max(0, Math.min(maxTokens, all.size())));
		for (int tokensToRemove = 0; tokensToRemove <= out.size(); tokensToRemove++) {
			final List<Integer> tokens = out.subList(0, out.size() - tokensToRemove);
			final String decoded = decode(tokens);
			if (text.startsWith(decoded)) return new EncodingResult(tokens, text.length() > decoded.length());
		}
		return new EncodingResult(out, false);
	}

	@Override
	public EncodingResult encode(final String text, final int maxTokens) {
		return encodeMax(text, maxTokens, false);
	}

	@Override
	public EncodingResult encodeOrdinary(final String text, final int maxTokens) {
		return encodeMax(text, maxTokens, true);
	}

	@Override
	public int countTokens(final String text) {
		if (text == null) return 0;
		final Batch b = encodeBatch(Collections.singletonList(text), false, true);
		raise(b.docStatus[0], text);
		return (int) b.tokenOffsets[1];
	}

	@Override
	public int countTokensOrdinary(final String text) {
		if (text == null) return 0;
		return (int) encodeBatch(Collections.singletonList(text), true, true).tokenOffsets[1];
	}

	@Override
	public byte[] decodeBytes(final List<Integer> tokens) {
		try (Arena arena = Arena.ofConfined()) {
			final MemorySegment ids = arena.allocate(JAVA_INT, Math.max(tokens.size(), 1)), off = arena.allocate(JAVA_LONG, 2);
			for (int i = 0; i < tokens.size(); i++) ids.setAtIndex(JAVA_INT, i, tokens.get(i));
			off.setAtIndex(JAVA_LONG, 0, 0L);
			off.setAtIndex(JAVA_LONG, 1, tokens.size());
			final MemorySegment out = arena.allocate(ADDRESS);
			final int rc = (int) JtkNative.DECODE_BATCH.invokeExact(handle, ids, off, 1L, out);
			if (rc != JtkNative.JTK_OK) throw new IllegalStateException(JtkNative.lastError());
			final MemorySegment r = out.get(ADDRESS, 0);
			try {
				final int st = ((MemorySegment) JtkNative.RESULT_DOC_STATUS.invokeExact(r)).reinterpret(4).get(JAVA_INT, 0);
				if ((st & JtkNative.DOC_UNKNOWN_ID) != 0) {
					final int bad = ((MemorySegment) JtkNative.RESULT_BAD_IDS.invokeExact(r)).reinterpret(4).get(JAVA_INT, 0);
					throw new IllegalArgumentException("Unknown token for decoding: " + bad); // GptBytePairEncoding.java:313
				}
				final long n = ((MemorySegment) JtkNative.RESULT_BYTE_OFFSETS.invokeExact(r)).reinterpret(16).getAtIndex(JAVA_LONG, 1);
				return ((MemorySegment) JtkNative.RESULT_BYTES.invokeExact(r)).reinterpret(n).toArray(JAVA_BYTE);
			} finally {
				JtkNative.RESULT_FREE.invokeExact(r);
			}
		} catch (final RuntimeException e) {
			throw e;
		} catch (final Throwable t) {
			throw new IllegalStateException(t);
		}
	}

	@Override
	public String decode(final List<Integer> tokens) {
		return new String(decodeBytes(tokens), StandardCharsets.UTF_8); // keeps the JVM's U+FFFD behaviour (:131-134)
	}

	@Override
	public String getName() {
		return name;
	}

	@Override
	public void close() {
		try {
			JtkNative.ENCODING_DESTROY.invokeExact(handle);
		} catch (final Throwable t) {
			throw new IllegalStateException(t);
		}
	}
}
