# Build recipe for the UNMODIFIED reference (`kat`, `kat_jellyfish`,
# `generate_sequence`) straight from the sources where they lie under
# $(REF) (default /root/reference).  Follows SURVEY.md §8(c): plain g++ on the
# reference's own .cc files, no autotools, no reference build system.
# Outputs go ONLY into oracle/_ref/ (git-ignored).  Nothing is copied into the
# repo.  Test infrastructure only: the product never links or runs this.
#
#   make -f oracle/ref.mk -j8          (from the repo root)

REF   ?= /root/reference
OUT   ?= oracle/_ref
OBJ   := $(OUT)/obj
J     := $(REF)/deps/jellyfish-2.2.0
BOOST := $(REF)/deps/boost

CXX      ?= g++
# -O2 -std=c++11 matches configure.ac:179; -fwrapv -m64 match lib/Makefile.am:48-52
CXXFLAGS := -std=c++11 -O2 -fwrapv -m64 -w -DHAVE_CONFIG_H -I$(OBJ) \
            -I$(J)/include -I$(J) -I$(BOOST) -I$(REF)/lib/include -I$(REF)/src \
            -I$(REF)/deps/seqan-library-2.0.0/include \
            -DKAT_SCRIPTS='"/tmp"' -DKAT_EXECPREFIX='"/tmp"' -DKAT_SITE_PKGS='"/tmp"'

JF_LIB_SRC := rectangular_binary_matrix.cc mer_dna.cc storage.cc allocators_mmap.cc \
              misc.cc int128.cc thread_exec.cc jsoncpp.cpp time.cc generator_manager.cc
JF_LIB_OBJ := $(addprefix $(OBJ)/jf_,$(addsuffix .o,$(basename $(JF_LIB_SRC))))

KATLIB_SRC := $(notdir $(wildcard $(REF)/lib/src/*.cc))
KATLIB_OBJ := $(addprefix $(OBJ)/katlib_,$(KATLIB_SRC:.cc=.o))
KAT_SRC    := $(notdir $(wildcard $(REF)/src/*.cc))
KAT_OBJ    := $(addprefix $(OBJ)/kat_,$(KAT_SRC:.cc=.o))

BOOST_LIBS := system filesystem chrono timer program_options
BOOST_SRC  := $(filter-out %windows_file_codecvt.cpp %winmain.cpp,\
              $(foreach l,$(BOOST_LIBS),$(wildcard $(BOOST)/libs/$(l)/src/*.cpp)))
BOOST_OBJ  := $(foreach s,$(BOOST_SRC),$(OBJ)/boost_$(word 2,$(subst /libs/, ,$(subst /src/,_,$(s)))).o)

JFCLI_SRC := $(notdir $(wildcard $(J)/sub_commands/*.cc))
JFCLI_OBJ := $(addprefix $(OBJ)/jfcli_,$(JFCLI_SRC:.cc=.o)) $(OBJ)/jfx_merge_files.o $(OBJ)/jfx_backtrace.o $(OBJ)/jfx_dbg.o
GENSEQ_OBJ := $(OBJ)/jfx_generate_sequence.o $(OBJ)/jfx_mersenne.o $(OBJ)/jfx_backtrace.o $(OBJ)/jfx_dbg.o

all: $(OUT)/kat $(OUT)/kat_jellyfish $(OUT)/generate_sequence

$(OBJ)/config.h:
	@mkdir -p $(OBJ)
	@printf '%s\n' '#define PACKAGE_NAME "kat"' '#define PACKAGE_VERSION "2.4.2"' \
	  '#define PACKAGE_STRING "kat 2.4.2"' '#define PACKAGE_URL ""' '#define PACKAGE_BUGREPORT ""' '#define VERSION "2.4.2"' \
	  '#define OS_LINUX 1' '#define HAVE_SSE 1' '#define HAVE_EXECINFO_H 1' \
	  '#define HAVE_EXT_STDIO_FILEBUF_H 1' '#define HAVE_CLOCK_GETTIME 1' '#define HAVE_SI_INT 1' > $@

$(OBJ)/jf_%.o: $(J)/lib/%.cc $(OBJ)/config.h
	$(CXX) $(CXXFLAGS) -c $< -o $@
$(OBJ)/jf_%.o: $(J)/lib/%.cpp $(OBJ)/config.h
	$(CXX) $(CXXFLAGS) -c $< -o $@
$(OBJ)/katlib_%.o: $(REF)/lib/src/%.cc $(OBJ)/config.h
	$(CXX) $(CXXFLAGS) -c $< -o $@
$(OBJ)/kat_%.o: $(REF)/src/%.cc $(OBJ)/config.h
	$(CXX) $(CXXFLAGS) -c $< -o $@
$(OBJ)/jfcli_%.o: $(J)/sub_commands/%.cc $(OBJ)/config.h
	$(CXX) $(CXXFLAGS) -c $< -o $@
$(OBJ)/jfx_%.o: $(J)/jellyfish/%.cc $(OBJ)/config.h
	$(CXX) $(CXXFLAGS) -c $< -o $@
$(OBJ)/jfx_%.o: $(J)/jellyfish/%.cpp $(OBJ)/config.h
	$(CXX) $(CXXFLAGS) -c $< -o $@

define BOOST_RULE
$(OBJ)/boost_$(word 2,$(subst /libs/, ,$(subst /src/,_,$(1)))).o: $(1)
	@mkdir -p $(OBJ)
	$(CXX) -std=c++11 -O2 -w -I$(BOOST) -c $$< -o $$@
endef
$(foreach s,$(BOOST_SRC),$(eval $(call BOOST_RULE,$(s))))

$(OUT)/kat: $(KAT_OBJ) $(KATLIB_OBJ) $(JF_LIB_OBJ) $(BOOST_OBJ)
	$(CXX) -o $@ $^ -lz -lpthread -lrt
$(OUT)/kat_jellyfish: $(JFCLI_OBJ) $(JF_LIB_OBJ)
	$(CXX) -o $@ $^ -lz -lpthread -lrt
$(OUT)/generate_sequence: $(GENSEQ_OBJ) $(JF_LIB_OBJ)
	$(CXX) -o $@ $^ -lz -lpthread -lrt

clean:
	rm -rf $(OUT)
.PHONY: all clean
