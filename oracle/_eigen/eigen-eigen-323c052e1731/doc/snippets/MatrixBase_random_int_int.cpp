cout << MatrixXi::Random(2,3) << endl;
