// Fabricate a format-exact Knlm model directory with the reference's OWN builder, because every binary
// model in the snapshot is a git-LFS pointer (models/cong/base/*.mdl, sj.morph) and models/base is absent.
// TEST INFRASTRUCTURE: links oracle/_ref/libkiwi_ref.so (the unmodified reference).
//   KiwiBuilder{ModelBuildArgs}                /root/reference/src/KiwiBuilder.cpp:1218-1250
//   KiwiBuilder::saveModel (sj.morph, sj.knlm)  KiwiBuilder.cpp:1479-1491
//   extract.mdl = two empty maps               src/WordDetector.cpp:175-191
// usage: fabricate_model <morphemes.txt> <outdir> <minMorphCnt> <corpus.txt>...
#include <fstream>
#include <iostream>
#include <map>
#include <kiwi/Kiwi.h>
#include "serializer.hpp"

using namespace kiwi;

int main(int argc, char** argv)
{
	if (argc < 5) { std::cerr << "usage: fabricate_model <morphemes.txt> <outdir> <minMorphCnt> <corpus>...\n"; return 2; }
	KiwiBuilder::ModelBuildArgs args;
	args.morphemeDef = argv[1];
	const std::string outDir = argv[2];
	args.minMorphCnt = std::stoul(argv[3]);
	for (int i = 4; i < argc; ++i) args.corpora.emplace_back(argv[i]);
	args.lmOrder = 4;
	args.lmMinCnts = { 1 };
	args.numWorkers = 1;
	args.useLmTagHistory = true;
	args.quantizeLm = true;
	args.compressLm = true;
	try
	{
		KiwiBuilder kb{ args };
		kb.saveModel(outDir);
		{
			std::ofstream ofs{ outDir + "/extract.mdl", std::ios_base::binary };
			std::map<std::pair<POSTag, bool>, std::map<char16_t, float>> posScore;
			std::map<std::u16string, float> nounTailScore;
			serializer::writeMany(ofs, posScore, nounTailScore);
		}
	}
	catch (const std::exception& e)
	{
		std::cerr << "fabricate_model failed: " << e.what() << std::endl;
		return 1;
	}
	return 0;
}
