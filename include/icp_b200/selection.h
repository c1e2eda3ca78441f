// selection.h -- drop-in for icp-variants/selection.h:8-107.  Inside estimatePose the selection (all / Bernoulli(p) per iteration) is a
// predicate evaluated by the device kernels (icp_gpu_config.selection / proba / seed); PointSelection is the same thing as the value class
// the reference's drivers can hold: a view of the whole cloud (SELECT_ALL) or the subset drawn by resample() with std::mt19937 +
// std::uniform_real_distribution<double>(0, 1) exactly as selection.h:88-104 draws it.  The reference seeds from std::random_device
// (:76-79); so does this class unless a seed is given (setSeed / the 4-argument constructor), which is what makes runs repeatable and
// lets the device loop (ICPOptimizer::setSelectionSeed) draw the same subsets.  One deliberate difference: resample() also clears the
// colours (the reference lets them pile up, :93-95).
#pragma once
#include <iostream>
#include <random>
#include <vector>
#include "PointCloud.h"

enum selection_methods { SELECT_ALL = 0, RANDOM_SAMPLING };
typedef std::mt19937 MyRNG;

class PointSelection {
public:
    PointSelection() : numPoints(0), numSelectedPoints(0), m_selectionMode(SELECT_ALL), m_selectionProba(0.5) {}
    PointSelection(const PointCloud& source, unsigned int selectionMode = SELECT_ALL, float selectionProba = 0.5f)
        : m_source{source}, numPoints((unsigned int)source.getPoints().size()), numSelectedPoints(0), m_selectionMode(selectionMode), m_selectionProba(selectionProba) {
        if (m_selectionMode > SELECT_ALL) { std::random_device rd; rng.seed(rd()); }
    }
    PointSelection(const PointCloud& source, unsigned int selectionMode, float selectionProba, unsigned int seed) : PointSelection(source, selectionMode, selectionProba) { rng.seed(seed); }
    void setSeed(unsigned int seed) { rng.seed(seed); }

    const std::vector<Vector3f>& getPoints() { return m_selectionMode == SELECT_ALL ? m_source.getPoints() : m_points; }
    const std::vector<Vector3f>& getNormals() { return m_selectionMode == SELECT_ALL ? m_source.getNormals() : m_normals; }
    const std::vector<Vector4uc>& getColors() { return m_selectionMode == SELECT_ALL ? m_source.getColors() : m_colors; }
    const std::vector<int>& getSelectedIndexes() const { return selectedPointIndexes; }      // extension: what icp_gpu_query_matches takes as sel_idx

    void resample() {
        std::cout << "Resample points.\n";
        std::uniform_real_distribution<double> ureal(0.0, 1.0);
        numSelectedPoints = 0;
        m_points.clear(); m_normals.clear(); m_colors.clear(); selectedPointIndexes.clear();
        for (size_t i = 0; i < numPoints; i++) {
            if (ureal(rng) < m_selectionProba) {
                m_points.push_back(m_source.getPoints()[i]);
                m_normals.push_back(m_source.getNormals()[i]);
                if (i < m_source.getColors().size()) m_colors.push_back(m_source.getColors()[i]);
                selectedPointIndexes.push_back((int)i);
                numSelectedPoints++;
            }
        }
        std::cout << "Number points samples " << m_points.size() << "\n";
    }

private:
    PointCloud m_source;
    std::vector<int> selectedPointIndexes;
    unsigned int numPoints, numSelectedPoints, m_selectionMode;
    double m_selectionProba;
    std::vector<Vector3f> m_points, m_normals;
    std::vector<Vector4uc> m_colors;
    MyRNG rng;
};
