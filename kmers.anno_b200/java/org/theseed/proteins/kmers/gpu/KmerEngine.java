package org.theseed.proteins.kmers.gpu;

import java.io.IOException;
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
 * ctypes binding kmers.anno_b200/engine.py and the C++ class host/KmerEngine.hpp exercise the
 * same C entry points one for one.  Java 21 is the reference's target (pom.xml:14-15), where
 * java.lang.foreign is still a preview API, hence JNI.
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
        byte[] packed = new byte[kmers.size() * this.kmerSize];
        int[] ids = new int[kmers.size()];
        for (int i = 0; i < ids.length; i++) {
            byte[] k = kmers.get(i).getBytes(StandardCharsets.ISO_8859_1);
            if (k.length != this.kmerSize)
                throw new IOException("Kmer database mixes kmer lengths at line " + (i + 1) + ".");
            System.arraycopy(k, 0, packed, i * this.kmerSize, this.kmerSize);
            final String role = roles.get(i);
            ids[i] = this.roleIds.computeIfAbsent(role, r -> { this.roleNames.add(r); return this.roleNames.size() - 1; });
        }
        dbLoad(this.handle, packed, ids, ids.length, this.kmerSize);
    }

    /** @return the role string of a dense id returned in {@link Calls#role} */
    public String getRole(int id) {
        return this.roleNames.get(id);
    }

    public int getKmerSize() {
        return this.kmerSize;
    }

    /**
     * Annotate a batch of proteins (the peg loop :122-148).  The caller replays
     * reporter.recordFeature(feat, getRole(role[i]), hits[i]) for every i with flag[i] == CALLED,
     * in the original peg order.
     */
    public Calls annotate(List<String> proteins, int minHits) throws IOException {
        long[] offsets = new long[proteins.size() + 1];
        int total = 0;
        for (int i = 0; i < proteins.size(); i++) { offsets[i] = total; total += proteins.get(i).length(); }
        offsets[proteins.size()] = total;
        byte[] residues = new byte[total];
        for (int i = 0; i < proteins.size(); i++) {
            byte[] p = proteins.get(i).getBytes(StandardCharsets.ISO_8859_1);
            System.arraycopy(p, 0, residues, (int) offsets[i], p.length);
        }
        Calls c = new Calls();
        c.role = new int[proteins.size()];
        c.hits = new int[proteins.size()];
        c.flag = new byte[proteins.size()];
        annotate(this.handle, residues, offsets, proteins.size(), minHits, c.role, c.hits, c.flag);
        return c;
    }

    /**
     * ProteinKmers.distance of every query protein against its candidates (GeneCopyProcessor.java:137-142).
     * Query q is proteins.get(querySeq[q]); its candidates are proteins.get(cand[m]) for
     * groupOff[q] <= m < groupOff[q+1].  Returns one distance per candidate entry; the caller keeps the
     * selection loop of :139-146.  No k-mer database is needed.
     */
    public double[] kmerDistance(List<String> proteins, int kmerSize, int[] querySeq, long[] groupOff, int[] cand)
            throws IOException {
        long[] offsets = new long[proteins.size() + 1];
        int total = 0;
        for (int i = 0; i < proteins.size(); i++) { offsets[i] = total; total += proteins.get(i).length(); }
        offsets[proteins.size()] = total;
        byte[] residues = new byte[total];
        for (int i = 0; i < proteins.size(); i++) {
            byte[] p = proteins.get(i).getBytes(StandardCharsets.ISO_8859_1);
            System.arraycopy(p, 0, residues, (int) offsets[i], p.length);
        }
        double[] dist = new double[cand.length];
        int[] common = new int[cand.length];
        kmerDistance(this.handle, residues, offsets, proteins.size(), kmerSize, querySeq, groupOff, querySeq.length,
                cand, common, dist);
        return dist;
    }

    @Override
    public void close() {
        if (this.handle != 0) { destroy(this.handle); this.handle = 0; }
    }

    // ---- native methods: each maps to one C-ABI entry point; a non-zero code becomes an IOException
    private static native long create(int[] devices) throws IOException;                     // ka_create
    private static native void destroy(long handle);                                          // ka_destroy
    private static native void dbLoad(long handle, byte[] kmers, int[] roleIds, long n, int k) throws IOException;  // ka_db_load
    private static native void annotate(long handle, byte[] residues, long[] offsets, long n, int minHits,
            int[] role, int[] hits, byte[] flag) throws IOException;                           // ka_annotate
    private static native void kmerDistance(long handle, byte[] residues, long[] offsets, long n, int k,
            int[] querySeq, long[] groupOff, long q, int[] cand, int[] common, double[] dist) throws IOException;  // ka_kmer_distance
}
