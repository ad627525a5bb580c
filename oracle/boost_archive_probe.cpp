// oracle/boost_archive_probe.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Writes DIR/database with the REAL Boost.Serialization library. This image has no Boost headers, but a header-less
// libboost_serialization.so (1.78) ships inside Nsight Compute. The classes below re-declare, with the library's own names
// and layouts, just enough of its interface to link against it: extended_type_info / basic_oserializer (the per-class
// serializer objects the header templates would instantiate), basic_oarchive::save_object, and text_oarchive_impl. What the
// library itself does here: the archive header (signature, library version), the class preamble (tracking, version) the
// first time each class is saved, token separation, std::string and item_version encoding. What is restated from the
// reference's serialize() methods (GenbankTools.h:57-62,100-109,154-163,197-200) and from Boost's traits defaults
// (object_class_info, track_selectively, version 0; vector = count, item_version, items): the order of the members.
// Input: the index dump format of oracle/ref_driver.cpp (kref_parsed_index_dump) extended by nothing; output: the archive.
//   usage: boost_archive_probe <dump file> <archive file>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <new>
#include <sstream>
#include <string>
#include <vector>

namespace boost {
namespace serialization {
class extended_type_info {
 private:
  virtual bool is_less_than(const extended_type_info &) const = 0;
  virtual bool is_equal(const extended_type_info &) const = 0;
  const unsigned int m_type_info_key;
  const char *m_key;
 protected:
  extended_type_info(const unsigned int type_info_key, const char *key);
  virtual ~extended_type_info();
 public:
  virtual const char *get_debug_info() const = 0;
  virtual void *construct(unsigned int = 0, ...) const = 0;
  virtual void destroy(void const *const) const = 0;
};
class item_version_type {
  unsigned int t;
 public:
  item_version_type() : t(0) {}
  explicit item_version_type(const unsigned int &t_) : t(t_) {}
  item_version_type(const item_version_type &o) : t(o.t) {}
};
}  // namespace serialization
namespace archive {
class version_type {
  uint_least32_t t;
 public:
  version_type() : t(0) {}
  explicit version_type(const unsigned int &t_) : t(t_) {}
  version_type(const version_type &o) : t(o.t) {}
};
namespace detail {
class basic_oarchive;
class basic_pointer_oserializer;
class basic_serializer {
  const boost::serialization::extended_type_info *m_eti;
 protected:
  explicit basic_serializer(const boost::serialization::extended_type_info &eti) : m_eti(&eti) {}
};
class basic_oserializer : public basic_serializer {
 private:
  basic_pointer_oserializer *m_bpos;
 protected:
  explicit basic_oserializer(const boost::serialization::extended_type_info &type_);
  virtual ~basic_oserializer();
 public:
  virtual void save_object_data(basic_oarchive &ar, const void *x) const = 0;
  virtual bool class_info() const = 0;
  virtual bool tracking(const unsigned int flags) const = 0;
  virtual version_type version() const = 0;
  virtual bool is_polymorphic() const = 0;
};
class basic_oarchive {
 public:
  void save_object(const void *x, const basic_oserializer &bos);
  void end_preamble();
};
}  // namespace detail
class text_oarchive;
template <class Archive> class basic_text_oarchive {
 public:
  void init();
  void newtoken();
};
template <class Archive> class text_oarchive_impl {
 public:
  text_oarchive_impl(std::ostream &os, unsigned int flags);
  ~text_oarchive_impl();
  void save(const std::string &s);
  void save(const boost::serialization::item_version_type &t);
};
}  // namespace archive
}  // namespace boost

using namespace boost::archive;
using boost::archive::detail::basic_oarchive;

struct Eti : boost::serialization::extended_type_info {
  Eti(unsigned key, const char *name) : extended_type_info(key, name) {}
  bool is_less_than(const extended_type_info &) const override { return false; }
  bool is_equal(const extended_type_info &) const override { return false; }
  const char *get_debug_info() const override { return "probe"; }
  void *construct(unsigned int, ...) const override { return nullptr; }
  void destroy(void const *const) const override {}
};

struct Ar {   // the archive object built by the library's own constructor in raw storage
  alignas(64) unsigned char raw[8192];
  std::ostream &os;
  explicit Ar(std::ostream &o) : os(o) {
    memset(raw, 0, sizeof raw);
    new (raw) text_oarchive_impl<text_oarchive>(o, 0);
    ((basic_text_oarchive<text_oarchive> *)raw)->init();
  }
  basic_oarchive *base() { return (basic_oarchive *)raw; }
  void token() { ((basic_text_oarchive<text_oarchive> *)raw)->newtoken(); }
  void str(const std::string &s) { base()->end_preamble(); ((text_oarchive_impl<text_oarchive> *)raw)->save(s); }
  template <class T> void num(T v) { base()->end_preamble(); token(); os << v; }
  void item_version(unsigned v) { ((text_oarchive_impl<text_oarchive> *)raw)->save(boost::serialization::item_version_type(v)); }
  void close() { ((text_oarchive_impl<text_oarchive> *)raw)->~text_oarchive_impl(); }   // the library's destructor ends the archive
};
static Ar *g_ar;

struct Ser : boost::archive::detail::basic_oserializer {
  void (*fn)(const void *);
  Ser(const Eti &e, void (*f)(const void *)) : basic_oserializer(e), fn(f) {}
  void save_object_data(basic_oarchive &, const void *x) const override { fn(x); }
  bool class_info() const override { return true; }              // implementation_level object_class_info
  bool tracking(const unsigned int) const override { return false; }   // track_selectively, never saved through a pointer
  version_type version() const override { return version_type(0); }
  bool is_polymorphic() const override { return false; }
};

struct CDS { uint32_t start, stop; bool complement; };
struct Gene { std::string s[5]; uint32_t geneID; CDS cds; };
struct Entry { std::string bases; uint32_t tax, gbid; bool plasmid, s16; std::string locus; std::vector<Gene> genes; };
struct Index { std::vector<Entry> entries; };

static Eti e_index(101, "Index"), e_entries(102, "Entries"), e_entry(103, "Entry"), e_genes(104, "Genes"), e_gene(105, "Gene"), e_cds(106, "CDS");
static void save_cds(const void *x);
static void save_gene(const void *x);
static void save_genes(const void *x);
static void save_entry(const void *x);
static void save_entries(const void *x);
static void save_index(const void *x);
static Ser s_index(e_index, save_index), s_entries(e_entries, save_entries), s_entry(e_entry, save_entry), s_genes(e_genes, save_genes),
    s_gene(e_gene, save_gene), s_cds(e_cds, save_cds);

static void save_cds(const void *x) { const CDS &c = *(const CDS *)x; g_ar->num(c.start); g_ar->num(c.stop); g_ar->num((int)c.complement); }
static void save_gene(const void *x) {
  const Gene &g = *(const Gene *)x;
  for (int i = 0; i < 5; i++) g_ar->str(g.s[i]);
  g_ar->num(g.geneID);
  g_ar->base()->save_object(&g.cds, s_cds);
}
static void save_genes(const void *x) {      // boost/serialization/vector.hpp -> collections_save_imp: count, item_version, items
  const std::vector<Gene> &v = *(const std::vector<Gene> *)x;
  g_ar->num((size_t)v.size());
  g_ar->item_version(0);
  for (auto &g : v) g_ar->base()->save_object(&g, s_gene);
}
static void save_entry(const void *x) {
  const Entry &e = *(const Entry *)x;
  g_ar->str(e.bases); g_ar->num(e.tax); g_ar->num(e.gbid); g_ar->num((int)e.plasmid); g_ar->num((int)e.s16); g_ar->str(e.locus);
  g_ar->base()->save_object(&e.genes, s_genes);
}
static void save_entries(const void *x) {
  const std::vector<Entry> &v = *(const std::vector<Entry> *)x;
  g_ar->num((size_t)v.size());
  g_ar->item_version(0);
  for (auto &e : v) g_ar->base()->save_object(&e, s_entry);
}
static void save_index(const void *x) { g_ar->base()->save_object(&((const Index *)x)->entries, s_entries); }

static std::vector<std::string> split(const std::string &s, char sep) {
  std::vector<std::string> out;
  size_t at = 0;
  for (;;) {
    const size_t p = s.find(sep, at);
    out.push_back(s.substr(at, p == std::string::npos ? std::string::npos : p - at));
    if (p == std::string::npos) break;
    at = p + 1;
  }
  return out;
}

int main(int argc, char **argv) {
  if (argc != 3) { fprintf(stderr, "usage: %s <dump> <archive>\n", argv[0]); return 2; }
  std::string data;
  { FILE *f = fopen(argv[1], "rb"); if (!f) return 2; char buf[65536]; size_t n; while ((n = fread(buf, 1, sizeof buf, f)) > 0) data.append(buf, n); fclose(f); }
  Index ix;
  for (const std::string &rec : split(data, 0x1e)) {
    if (rec.empty()) continue;
    const std::vector<std::string> f = split(rec, 0x1f);
    if (f[0] == "E" && f.size() == 7) {
      Entry e;
      e.locus = f[1]; e.tax = (uint32_t)std::stoul(f[2]); e.gbid = (uint32_t)std::stoul(f[3]); e.plasmid = f[4] != "0"; e.s16 = f[5] != "0"; e.bases = f[6];
      ix.entries.push_back(e);
    } else if (f[0] == "G" && f.size() == 10 && !ix.entries.empty()) {
      Gene g;
      for (int i = 0; i < 5; i++) g.s[i] = f[1 + i];
      g.geneID = (uint32_t)std::stoul(f[6]); g.cds.start = (uint32_t)std::stoul(f[7]); g.cds.stop = (uint32_t)std::stoul(f[8]); g.cds.complement = f[9] != "0";
      ix.entries.back().genes.push_back(g);
    } else { fprintf(stderr, "bad record\n"); return 2; }
  }
  std::ostringstream os;
  {
    Ar ar(os);
    g_ar = &ar;
    ar.base()->save_object(&ix, s_index);
    ar.close();
  }
  FILE *o = fopen(argv[2], "wb");
  if (!o) return 2;
  fwrite(os.str().data(), 1, os.str().size(), o);
  fclose(o);
  return 0;
}
