/* TEST INFRASTRUCTURE ONLY — minimal stand-in for the HTSlib 1.17 VCF API.
 *
 * HTSlib (pinned at v1.17 by the reference's README.md:86-88, src/Dockerfile:25)
 * is not installed in this image and cannot be built offline.  It is used by the
 * reference only for I/O (src/variant.cpp:413-995, src/globals.cpp:80-92); the
 * hot path (src/dist.cpp) never touches it.  This header implements just the
 * subset of the API those call sites use, with the semantics that matter for
 * them, so that the UNMODIFIED reference sources compile into oracle/_ref/.
 *
 * Written from the documented behaviour of the API (not copied from HTSlib):
 *   - text VCF, plain / gzip / bgzip, read through zlib's gzFile
 *   - one sample column
 *   - header records for FILTER and contig lines carry an added "IDX" key
 *     (PASS is always IDX 0)
 *   - bcf_get_format_int32/float: -1 tag not in header, -2 type clash,
 *     -3 tag absent from the record, else the number of values
 *   - GT ints are (allele+1)<<1 | phased, first allele's phase bit always 0,
 *     '.' allele encodes as 0
 *   - missing QUAL is NaN, pos is 0-based
 */
#ifndef VD_STANDIN_HTSLIB_VCF_H
#define VD_STANDIN_HTSLIB_VCF_H

#include <zlib.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <vector>

#define BCF_HL_FLT  0
#define BCF_HL_INFO 1
#define BCF_HL_FMT  2
#define BCF_HL_CTG  3
#define BCF_HL_STR  4
#define BCF_HL_GEN  5
#define BCF_UN_ALL  15

#define bcf_int32_missing     (INT32_MIN)
#define bcf_int32_vector_end  (INT32_MIN + 1)

#define bcf_gt_phased(idx)    (((idx) + 1) << 1 | 1)
#define bcf_gt_unphased(idx)  (((idx) + 1) << 1)
#define bcf_gt_missing        0
#define bcf_gt_is_missing(val) ((val) >> 1 ? 0 : 1)
#define bcf_gt_is_phased(idx)  ((idx) & 1)
#define bcf_gt_allele(val)     (((val) >> 1) - 1)

typedef struct {
    int type;
    char *key;
    char *value;
    int nkeys;
    char **keys, **vals;
} bcf_hrec_t;

typedef struct {
    int nhrec;
    bcf_hrec_t **hrec;
    char **samples;
    /* stand-in private state */
    int n_samples_;
    std::vector<std::string> *ctg_names_;
    std::map<std::string, int> *ctg_ids_;
    std::map<std::string, int> *flt_ids_;
    std::map<std::string, int> *fmt_types_;   /* 0 int, 1 float, 2 string, 3 other */
} bcf_hdr_t;

typedef struct {
    int n_flt;
    int *flt;
    char **allele;
    /* stand-in private */
    int n_allele_;
} bcf_dec_t;

typedef struct {
    int64_t pos;
    int32_t rid;
    float qual;
    bcf_dec_t d;
    /* stand-in private: FORMAT keys and sample values of the current line */
    std::vector<std::string> *fmt_keys_;
    std::vector<std::string> *smp_vals_;
    std::vector<std::string> *allele_store_;
    std::vector<int> *flt_store_;
} bcf1_t;

typedef struct {
    gzFile fp;
    std::string *pending;   /* first non-header line, read ahead by bcf_hdr_read */
    bool has_pending;
} htsFile;

/* ---------------------------------------------------------------------- */

static inline bool vds_getline_(gzFile fp, std::string &line) {
    line.clear();
    char buf[65536];
    bool any = false;
    while (gzgets(fp, buf, sizeof(buf)) != NULL) {
        any = true;
        size_t n = strlen(buf);
        if (n && buf[n - 1] == '\n') {
            line.append(buf, n - 1);
            if (!line.empty() && line.back() == '\r') line.pop_back();
            return true;
        }
        line.append(buf, n);
    }
    return any;
}

static inline std::vector<std::string> vds_split_(const std::string &s, char sep) {
    std::vector<std::string> out;
    size_t b = 0;
    while (true) {
        size_t e = s.find(sep, b);
        if (e == std::string::npos) { out.push_back(s.substr(b)); break; }
        out.push_back(s.substr(b, e - b));
        b = e + 1;
    }
    return out;
}

static inline htsFile *bcf_open(const char *fn, const char *mode) {
    (void)mode;
    gzFile fp = gzopen(fn, "rb");
    if (fp == NULL) return NULL;
    htsFile *f = new htsFile;
    f->fp = fp;
    f->pending = new std::string();
    f->has_pending = false;
    return f;
}

static inline int bcf_close(htsFile *f) {
    if (!f) return -1;
    gzclose(f->fp);
    delete f->pending;
    delete f;
    return 0;
}

static inline char *vds_strdup_(const std::string &s) {
    char *p = (char *)malloc(s.size() + 1);
    memcpy(p, s.data(), s.size() + 1);
    return p;
}

/* parse "<k1=v1,k2="quoted, v",...>" into key/value lists */
static inline void vds_parse_struct_(const std::string &body,
        std::vector<std::string> &keys, std::vector<std::string> &vals) {
    size_t i = 0, n = body.size();
    while (i < n) {
        size_t eq = body.find('=', i);
        if (eq == std::string::npos) break;
        std::string key = body.substr(i, eq - i);
        std::string val;
        size_t j = eq + 1;
        if (j < n && body[j] == '"') {
            j++;
            while (j < n && body[j] != '"') {
                if (body[j] == '\\' && j + 1 < n) { val.push_back(body[j + 1]); j += 2; }
                else val.push_back(body[j++]);
            }
            j++; /* closing quote */
            while (j < n && body[j] != ',') j++;
        } else {
            while (j < n && body[j] != ',') val.push_back(body[j++]);
        }
        keys.push_back(key);
        vals.push_back(val);
        i = j + 1;
    }
}

static inline void vds_add_hrec_(bcf_hdr_t *h, int type, const std::string &key,
        const std::string &value, std::vector<std::string> keys,
        std::vector<std::string> vals) {
    bcf_hrec_t *r = (bcf_hrec_t *)calloc(1, sizeof(bcf_hrec_t));
    r->type = type;
    r->key = vds_strdup_(key);
    r->value = value.empty() ? NULL : vds_strdup_(value);
    r->nkeys = (int)keys.size();
    r->keys = (char **)calloc(keys.size() + 1, sizeof(char *));
    r->vals = (char **)calloc(vals.size() + 1, sizeof(char *));
    for (size_t i = 0; i < keys.size(); i++) {
        r->keys[i] = vds_strdup_(keys[i]);
        r->vals[i] = vds_strdup_(vals[i]);
    }
    h->hrec = (bcf_hrec_t **)realloc(h->hrec, sizeof(bcf_hrec_t *) * (h->nhrec + 1));
    h->hrec[h->nhrec++] = r;
}

static inline bcf_hdr_t *bcf_hdr_read(htsFile *f) {
    bcf_hdr_t *h = new bcf_hdr_t;
    h->nhrec = 0;
    h->hrec = NULL;
    h->samples = NULL;
    h->n_samples_ = 0;
    h->ctg_names_ = new std::vector<std::string>();
    h->ctg_ids_ = new std::map<std::string, int>();
    h->flt_ids_ = new std::map<std::string, int>();
    h->fmt_types_ = new std::map<std::string, int>();

    /* PASS always exists with IDX 0 */
    (*h->flt_ids_)["PASS"] = 0;
    bool pass_emitted = false;

    std::string line;
    while (vds_getline_(f->fp, line)) {
        if (line.rfind("##", 0) == 0) {
            size_t eq = line.find('=');
            if (eq == std::string::npos) continue;
            std::string key = line.substr(2, eq - 2);
            std::string rest = line.substr(eq + 1);
            if (!rest.empty() && rest[0] == '<' && rest.back() == '>') {
                std::vector<std::string> keys, vals;
                vds_parse_struct_(rest.substr(1, rest.size() - 2), keys, vals);
                std::string id;
                for (size_t i = 0; i < keys.size(); i++) if (keys[i] == "ID") id = vals[i];
                int type = BCF_HL_STR;
                if (key == "FILTER") {
                    type = BCF_HL_FLT;
                    int idx;
                    if (h->flt_ids_->count(id)) idx = (*h->flt_ids_)[id];
                    else { idx = (int)h->flt_ids_->size(); (*h->flt_ids_)[id] = idx; }
                    if (id == "PASS") pass_emitted = true;
                    keys.push_back("IDX"); vals.push_back(std::to_string(idx));
                } else if (key == "contig") {
                    type = BCF_HL_CTG;
                    int idx = (int)h->ctg_names_->size();
                    if (!h->ctg_ids_->count(id)) {
                        (*h->ctg_ids_)[id] = idx;
                        h->ctg_names_->push_back(id);
                    } else idx = (*h->ctg_ids_)[id];
                    keys.push_back("IDX"); vals.push_back(std::to_string(idx));
                } else if (key == "FORMAT") {
                    type = BCF_HL_FMT;
                    std::string ty;
                    for (size_t i = 0; i < keys.size(); i++) if (keys[i] == "Type") ty = vals[i];
                    int t = ty == "Integer" ? 0 : ty == "Float" ? 1 : ty == "String" ? 2 : 3;
                    (*h->fmt_types_)[id] = t;
                } else if (key == "INFO") {
                    type = BCF_HL_INFO;
                }
                vds_add_hrec_(h, type, key, "", keys, vals);
            } else {
                vds_add_hrec_(h, BCF_HL_GEN, key, rest, {}, {});
            }
        } else if (line.rfind("#CHROM", 0) == 0) {
            std::vector<std::string> cols = vds_split_(line, '\t');
            int ns = cols.size() > 9 ? (int)cols.size() - 9 : 0;
            h->n_samples_ = ns;
            h->samples = (char **)calloc(ns + 1, sizeof(char *));
            for (int i = 0; i < ns; i++) h->samples[i] = vds_strdup_(cols[9 + i]);
            break;
        } else {
            /* no #CHROM line: keep the line for the record reader */
            *f->pending = line;
            f->has_pending = true;
            break;
        }
    }
    if (!pass_emitted) {
        vds_add_hrec_(h, BCF_HL_FLT, "FILTER", "",
                {"ID", "Description", "IDX"}, {"PASS", "All filters passed", "0"});
    }
    return h;
}

static inline int bcf_hdr_nsamples(const bcf_hdr_t *h) { return h->n_samples_; }

static inline const char **bcf_hdr_seqnames(const bcf_hdr_t *h, int *n) {
    *n = (int)h->ctg_names_->size();
    const char **names = (const char **)calloc(*n + 1, sizeof(char *));
    for (int i = 0; i < *n; i++) names[i] = (*h->ctg_names_)[i].c_str();
    return names;
}

static inline void bcf_hdr_destroy(bcf_hdr_t *h) {
    if (!h) return;
    for (int i = 0; i < h->nhrec; i++) {
        bcf_hrec_t *r = h->hrec[i];
        free(r->key); free(r->value);
        for (int j = 0; j < r->nkeys; j++) { free(r->keys[j]); free(r->vals[j]); }
        free(r->keys); free(r->vals); free(r);
    }
    free(h->hrec);
    for (int i = 0; i < h->n_samples_; i++) free(h->samples[i]);
    free(h->samples);
    delete h->ctg_names_; delete h->ctg_ids_; delete h->flt_ids_; delete h->fmt_types_;
    delete h;
}

static inline bcf1_t *bcf_init() {
    bcf1_t *r = new bcf1_t;
    r->pos = 0; r->rid = -1; r->qual = 0;
    r->d.n_flt = 0; r->d.flt = NULL; r->d.allele = NULL; r->d.n_allele_ = 0;
    r->fmt_keys_ = new std::vector<std::string>();
    r->smp_vals_ = new std::vector<std::string>();
    r->allele_store_ = new std::vector<std::string>();
    r->flt_store_ = new std::vector<int>();
    return r;
}

static inline void bcf_destroy(bcf1_t *r) {
    if (!r) return;
    free(r->d.allele);
    delete r->fmt_keys_; delete r->smp_vals_; delete r->allele_store_; delete r->flt_store_;
    delete r;
}

/* returns 0 on success, -1 on EOF */
static inline int bcf_read(htsFile *f, bcf_hdr_t *h, bcf1_t *r) {
    std::string line;
    while (true) {
        if (f->has_pending) { line = *f->pending; f->has_pending = false; }
        else if (!vds_getline_(f->fp, line)) return -1;
        if (line.empty() || line[0] == '#') continue;
        break;
    }
    std::vector<std::string> c = vds_split_(line, '\t');
    if (c.size() < 8) return -2;

    /* CHROM: contigs absent from the header are appended, as htslib does (with a warning) */
    auto it = h->ctg_ids_->find(c[0]);
    if (it == h->ctg_ids_->end()) {
        int idx = (int)h->ctg_names_->size();
        (*h->ctg_ids_)[c[0]] = idx;
        h->ctg_names_->push_back(c[0]);
        r->rid = idx;
    } else r->rid = it->second;

    r->pos = strtoll(c[1].c_str(), NULL, 10) - 1;
    r->qual = (c[5] == ".") ? std::numeric_limits<float>::quiet_NaN()
                            : strtof(c[5].c_str(), NULL);

    r->allele_store_->clear();
    r->allele_store_->push_back(c[3]);
    if (c[4] != ".") {
        std::vector<std::string> alts = vds_split_(c[4], ',');
        for (auto &a : alts) r->allele_store_->push_back(a);
    }
    r->d.n_allele_ = (int)r->allele_store_->size();
    r->d.allele = (char **)realloc(r->d.allele, sizeof(char *) * (r->d.n_allele_ + 1));
    for (int i = 0; i < r->d.n_allele_; i++)
        r->d.allele[i] = (char *)(*r->allele_store_)[i].c_str();

    r->flt_store_->clear();
    if (c[6] != ".") {
        std::vector<std::string> fl = vds_split_(c[6], ';');
        for (auto &name : fl) {
            auto fi = h->flt_ids_->find(name);
            int idx;
            if (fi == h->flt_ids_->end()) {
                idx = (int)h->flt_ids_->size();
                (*h->flt_ids_)[name] = idx;
            } else idx = fi->second;
            r->flt_store_->push_back(idx);
        }
    }
    r->d.n_flt = (int)r->flt_store_->size();
    r->d.flt = r->d.n_flt ? r->flt_store_->data() : NULL;

    r->fmt_keys_->clear();
    r->smp_vals_->clear();
    if (c.size() >= 10) {
        *r->fmt_keys_ = vds_split_(c[8], ':');
        *r->smp_vals_ = vds_split_(c[9], ':');
    }
    return 0;
}

static inline int bcf_unpack(bcf1_t *r, int which) { (void)r; (void)which; return 0; }

static inline int vds_find_fmt_(bcf1_t *r, const char *tag, std::string &val) {
    for (size_t i = 0; i < r->fmt_keys_->size(); i++) {
        if ((*r->fmt_keys_)[i] == tag) {
            if (i >= r->smp_vals_->size()) return 0;   /* trailing fields dropped */
            val = (*r->smp_vals_)[i];
            return 1;
        }
    }
    return 0;
}

static inline int bcf_get_format_int32(const bcf_hdr_t *h, bcf1_t *r, const char *tag,
        int32_t **dst, int *ndst) {
    bool is_gt = strcmp(tag, "GT") == 0;
    auto t = h->fmt_types_->find(tag);
    if (t == h->fmt_types_->end()) return -1;
    if (!is_gt && t->second != 0) return -2;
    if (is_gt && t->second != 2) return -2;
    std::string val;
    if (!vds_find_fmt_(r, tag, val)) return -3;

    std::vector<int32_t> out;
    if (is_gt) {
        size_t i = 0, n = val.size();
        bool phased = false;   /* first allele: phase bit 0 */
        while (i <= n) {
            size_t j = i;
            while (j < n && val[j] != '|' && val[j] != '/') j++;
            std::string a = val.substr(i, j - i);
            int32_t enc;
            if (a == "." || a.empty()) enc = 0 | (phased ? 1 : 0);
            else enc = ((atoi(a.c_str()) + 1) << 1) | (phased ? 1 : 0);
            out.push_back(enc);
            if (j >= n) break;
            phased = (val[j] == '|');
            i = j + 1;
        }
    } else {
        std::vector<std::string> parts = vds_split_(val, ',');
        for (auto &p : parts)
            out.push_back(p == "." ? bcf_int32_missing : (int32_t)strtol(p.c_str(), NULL, 10));
    }
    int n = (int)out.size();
    if (*ndst < n || *dst == NULL) {
        *dst = (int32_t *)realloc(*dst, sizeof(int32_t) * n);
        *ndst = n;
    }
    memcpy(*dst, out.data(), sizeof(int32_t) * n);
    return n;
}

static inline int bcf_get_format_float(const bcf_hdr_t *h, bcf1_t *r, const char *tag,
        float **dst, int *ndst) {
    auto t = h->fmt_types_->find(tag);
    if (t == h->fmt_types_->end()) return -1;
    if (t->second != 1) return -2;
    std::string val;
    if (!vds_find_fmt_(r, tag, val)) return -3;
    std::vector<std::string> parts = vds_split_(val, ',');
    int n = (int)parts.size();
    if (*ndst < n || *dst == NULL) {
        *dst = (float *)realloc(*dst, sizeof(float) * n);
        *ndst = n;
    }
    for (int i = 0; i < n; i++)
        (*dst)[i] = parts[i] == "." ? std::numeric_limits<float>::quiet_NaN()
                                     : strtof(parts[i].c_str(), NULL);
    return n;
}

#endif
