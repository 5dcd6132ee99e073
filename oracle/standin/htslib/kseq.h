/* TEST INFRASTRUCTURE ONLY — minimal stand-in for HTSlib's kseq.h FASTA reader,
 * covering the four calls the reference makes (src/fasta.h:10-23):
 * KSEQ_INIT(int, read), kseq_init(fd), kseq_read(seq), kseq_destroy(seq), and
 * the fields seq->name.s / seq->seq.s.  Plain (uncompressed) FASTA/FASTQ-less
 * input only; the record name is the header up to the first whitespace.
 * Written from the API's documented behaviour, not copied from HTSlib.
 */
#ifndef VD_STANDIN_HTSLIB_KSEQ_H
#define VD_STANDIN_HTSLIB_KSEQ_H

#include <unistd.h>
#include <cstdlib>
#include <cstring>
#include <string>

typedef struct { size_t l, m; char *s; } kstring_t;

typedef struct {
    kstring_t name, comment, seq, qual;
    /* stand-in private */
    int fd_;
    char *buf_;
    int beg_, end_, eof_;
    std::string *name_, *seq_;
} kseq_t;

#define VDS_KS_BUFSZ (1 << 20)

static inline int vds_ks_getc_(kseq_t *k) {
    if (k->beg_ >= k->end_) {
        if (k->eof_) return -1;
        k->end_ = (int)read(k->fd_, k->buf_, VDS_KS_BUFSZ);
        k->beg_ = 0;
        if (k->end_ <= 0) { k->eof_ = 1; return -1; }
    }
    return (unsigned char)k->buf_[k->beg_++];
}

static inline kseq_t *vds_kseq_init_(int fd) {
    kseq_t *k = (kseq_t *)calloc(1, sizeof(kseq_t));
    k->fd_ = fd;
    k->buf_ = (char *)malloc(VDS_KS_BUFSZ);
    k->name_ = new std::string();
    k->seq_ = new std::string();
    return k;
}

static inline void vds_kseq_destroy_(kseq_t *k) {
    if (!k) return;
    free(k->buf_);
    delete k->name_;
    delete k->seq_;
    free(k);
}

/* returns sequence length (>= 0) or -1 at EOF */
static inline int vds_kseq_read_(kseq_t *k) {
    int c;
    /* find next '>' */
    while ((c = vds_ks_getc_(k)) >= 0 && c != '>') {}
    if (c < 0) return -1;
    k->name_->clear();
    k->seq_->clear();
    /* name: up to whitespace */
    while ((c = vds_ks_getc_(k)) >= 0 && c != '\n' && c != ' ' && c != '\t' && c != '\r')
        k->name_->push_back((char)c);
    /* rest of header line */
    while (c >= 0 && c != '\n') c = vds_ks_getc_(k);
    /* sequence lines until next '>' at line start */
    bool line_start = true;
    while (true) {
        /* peek */
        if (k->beg_ >= k->end_) {
            if (k->eof_) break;
            k->end_ = (int)read(k->fd_, k->buf_, VDS_KS_BUFSZ);
            k->beg_ = 0;
            if (k->end_ <= 0) { k->eof_ = 1; break; }
        }
        /* bulk-scan the buffer */
        while (k->beg_ < k->end_) {
            char ch = k->buf_[k->beg_];
            if (line_start && ch == '>') goto done;
            k->beg_++;
            if (ch == '\n') { line_start = true; continue; }
            line_start = false;
            if (ch == '\r' || ch == ' ' || ch == '\t') continue;
            k->seq_->push_back(ch);
        }
    }
done:
    k->name.s = (char *)k->name_->c_str(); k->name.l = k->name_->size();
    k->seq.s = (char *)k->seq_->c_str();   k->seq.l = k->seq_->size();
    return (int)k->seq_->size();
}

#define KSEQ_INIT(type_t, __read)                                            \
    static inline kseq_t *kseq_init(type_t fd) { return vds_kseq_init_(fd); } \
    static inline int kseq_read(kseq_t *k) { return vds_kseq_read_(k); }      \
    static inline void kseq_destroy(kseq_t *k) { vds_kseq_destroy_(k); }

#endif
