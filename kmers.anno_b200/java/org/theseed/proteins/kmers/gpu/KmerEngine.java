package org.theseed.proteins.kmers.gpu;

import java.io.IOException;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.nio.charset.StandardCharsets;
import java.util.ArrayList;
import java.util.HashMap;
import java.util.List;
import java.util.Map;

/**
 * JNI face of libkmeranno.so (include/kmeranno.h).  One instance replaces the
 * {@code Map<String, String> kmerRoleMap} of ApplyKmerProcessor (ApplyKmerProcessor.java:53)
 * and runs its peg loop (:122-148) for a whole batch of proteins on the GPU(s).
 *
 * NOT COMPILED IN THIS REPOSITORY: the authoring image has no JDK (SURVEY.md fact 4).  The
 * ctypes binding kmers.anno_b200/engine.py and the C++ classes host/KmerEngine.hpp and
 * host/PackedBatch.* exercise the same C entry points one for one; the JNI shim is
 * syntax-checked against a stub jni.h (tests/test_host.py).  Java 21 is the reference's target
 * (pom.xml:14-15), where java.lang.foreign is still a preview API, hence JNI.
 *
 * All bulk data crosses in direct ByteBuffers over PINNED host memory ({@link #allocPinned}):
 * the Java side writes the batch once, in the engine's packed input form (5-bit residue codes +
 * 32-bit offsets, 0.625 bytes per residue over PCIe), while it touches the protein strings anyway.
 */
public final class KmerEngine implements AutoCloseable {

    static {
        System.loadLibrary("kmerengine_jni");   // jni/kmerengine_jni.c, linked against libkmeranno.so
    }

    /** per-protein outcome codes, KA_FLAG_* */
    public static final byte NONE = 0, CALLED = 1, AMBIGUOUS = 2, BELOW_MIN = 3;

    /** result of one annotate call: parallel arrays, one entry per protein, in input order */
    public static final class Calls {
        public int[] role;   // dense role id, -1 = no call
        public int[] hits;   // distinct hitting kmers when unanimous
        public byte[] flag;
    }

    private long handle;                       // ka_engine*
    private final List<String> roleNames = new ArrayList<>();
    private final Map<String, Integer> roleIds = new HashMap<>();
    private int kmerSize;
    private byte[] codeOfByte;                 // ka_db_get_alphabet: 0..n-1 for the DB's residues, 31 otherwise
    // pinned buffers of the current batch, grown on demand
    private ByteBuffer codes, offsets, role, hits, flag;

    public KmerEngine(int[] devices) throws IOException {
        this.handle = create(devices);
    }

    /**
     * Load the kmer database: the replacement of the put() loop of ApplyKmerProcessor.java:102-107.
     * kmers are the first column of kmerdb.tbl, roles the second; the LAST line of a repeated
     * kmer wins, as HashMap.put does.  All kmers must have the same length (the Java map does
     * not care, the packed table does).
     */
    public void loadDb(List<String> kmers, List<String> roles) throws IOException {
        if (kmers.isEmpty()) throw new IOException("Empty kmer database.");
        this.kmerSize = kmers.get(kmers.size() - 1).length();          // :108
        final int n = kmers.size();
        ByteBuffer km = allocPinned((long) n * this.kmerSize);
        ByteBuffer ids = allocPinned((long) n * 4).order(ByteOrder.LITTLE_ENDIAN);
        try {
            for (int i = 0; i < n; i++) {
                byte[] k = kmers.get(i).getBytes(StandardCharsets.ISO_8859_1);
                if (k.length != this.kmerSize)
                    throw new IOException("Kmer database mixes kmer lengths at line " + (i + 1) + ".");
                km.put(k);
                final String r = roles.get(i);
                ids.putInt(this.roleIds.computeIfAbsent(r, x -> { this.roleNames.add(x); return this.roleNames.size() - 1; }));
            }
            dbLoad(this.handle, km, ids, n, this.kmerSize);
            this.codeOfByte = alphabet(this.handle);
        } finally {
            freePinned(km);
            freePinned(ids);
        }
    }

    /** @return the role string of a dense id returned in {@link Calls#role} */
    public String getRole(int id) {
        return this.roleNames.get(id);
    }

    public int getKmerSize() {
        return this.kmerSize;
    }

    /**
     * Annotate a batch of proteins (the peg loop :122-148).  The proteins are written once, as 5-bit
     * codes, into the pinned stream (residue r of the batch at bits [5r, 5r+5), little endian).  The
     * caller replays reporter.recordFeature(feat, getRole(role[i]), hits[i]) for every i with
     * flag[i] == CALLED, in the original peg order.
     */
    public Calls annotate(List<String> proteins, int minHits) throws IOException {
        final int n = proteins.size();
        long total = 0;
        for (String p : proteins) total += p.length();
        if (total >= (1L << 32)) throw new IOException("More than 2^32 residues in one batch.");
        reserve(total, n);
        long acc = 0;          // bits not yet written
        int nbits = 0;
        long r = 0;
        this.codes.clear();
        this.offsets.clear();
        for (String p : proteins) {
            this.offsets.putInt((int) r);
            for (int i = 0; i < p.length(); i++) {
                final char ch = p.charAt(i);
                final long c = ch < 256 ? (this.codeOfByte[ch] & 31) : 31;   // a char outside Latin-1 is in no kmer
                acc |= c << nbits;
                nbits += 5;
                if (nbits >= 40) { putBits40(acc); acc >>>= 40; nbits -= 40; }
            }
            r += p.length();
        }
        this.offsets.putInt((int) r);
        for (; nbits > 0; nbits -= 8, acc >>>= 8) this.codes.put((byte) acc);
        annotatePacked(this.handle, this.codes, (total * 5 + 7) / 8, this.offsets, n, minHits, this.role, this.hits, this.flag);
        Calls c = new Calls();
        c.role = new int[n];
        c.hits = new int[n];
        c.flag = new byte[n];
        this.role.clear(); this.role.asIntBuffer().get(c.role, 0, n);
        this.hits.clear(); this.hits.asIntBuffer().get(c.hits, 0, n);
        this.flag.clear(); this.flag.get(c.flag, 0, n);
        return c;
    }

    private void putBits40(long v) {
        this.codes.put((byte) v).put((byte) (v >>> 8)).put((byte) (v >>> 16)).put((byte) (v >>> 24)).put((byte) (v >>> 32));
    }

    private void reserve(long residues, int n) throws IOException {
        final long needCodes = (residues * 5 + 7) / 8 + 64;
        if (this.codes == null || this.codes.capacity() < needCodes) {
            freePinned(this.codes);
            this.codes = allocPinned(needCodes + needCodes / 4);
        }
        if (this.offsets == null || this.offsets.capacity() < 4L * (n + 1)) {
            freePinned(this.offsets); freePinned(this.role); freePinned(this.hits); freePinned(this.flag);
            final long cap = n + 1 + n / 4;
            this.offsets = allocPinned(4 * cap).order(ByteOrder.LITTLE_ENDIAN);
            this.role = allocPinned(4 * cap).order(ByteOrder.LITTLE_ENDIAN);
            this.hits = allocPinned(4 * cap).order(ByteOrder.LITTLE_ENDIAN);
            this.flag = allocPinned(cap);
        }
    }

    @Override
    public void close() {
        freePinned(this.codes); freePinned(this.offsets); freePinned(this.role); freePinned(this.hits); freePinned(this.flag);
        this.codes = this.offsets = this.role = this.hits = this.flag = null;
        if (this.handle != 0) { destroy(this.handle); this.handle = 0; }
    }

    // ---- native methods: each maps to one C-ABI entry point; a non-zero code becomes an IOException
    private static native long create(int[] devices) throws IOException;                     // ka_create
    private static native void destroy(long handle);                                          // ka_destroy
    public static native ByteBuffer allocPinned(long bytes) throws IOException;               // ka_host_alloc
    public static native void freePinned(ByteBuffer buffer);                                  // ka_host_free (null is fine)
    private static native void dbLoad(long handle, ByteBuffer kmers, ByteBuffer roleIds, long n, int k) throws IOException;  // ka_db_load
    private static native byte[] alphabet(long handle) throws IOException;                    // ka_db_get_alphabet
    private static native void annotatePacked(long handle, ByteBuffer codes, long codeBytes, ByteBuffer offsets, long n,
            int minHits, ByteBuffer role, ByteBuffer hits, ByteBuffer flag) throws IOException;       // ka_annotate_packed
    public static native void annotate(long handle, ByteBuffer residues, long residueBytes, ByteBuffer offsets, long n,
            int minHits, ByteBuffer role, ByteBuffer hits, ByteBuffer flag) throws IOException;       // ka_annotate
    public static native void kmerDistance(long handle, ByteBuffer residues, long residueBytes, ByteBuffer offsets, long n, int k,
            ByteBuffer querySeq, ByteBuffer groupOff, long q, ByteBuffer cand, long m, ByteBuffer common, ByteBuffer dist)
            throws IOException;                                                                // ka_kmer_distance
}
