#define HTS_VERSION "1.3.1-genomicsdb-oracle"
